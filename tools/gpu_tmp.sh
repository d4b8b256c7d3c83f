#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/xblock_check.py check 2>&1 | grep "XB check" | cut -c1-120
timeout 300 python tools/xblock_check.py timeline 2>&1 | grep "per-CTA" | head -1
timeout 300 python tools/xblock_check.py bench 2>&1 | grep "XB bench"
UNIMP_XB_FLAGS=3 timeout 300 python tools/xblock_check.py bench 2>&1 | grep "XB bench"
