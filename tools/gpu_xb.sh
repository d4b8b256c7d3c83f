#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2x}
timeout 120 python tools/xblock_check.py check 2>&1 | grep -E "XB check|unimp|Error" | head -8
timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline.log 2>&1
grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline.log | head -4
timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"
UNIMP_XB_FLAGS=0 timeout 120 python tools/xblock_check.py timeline 2>&1 | grep -E "per-CTA durations" | sed "s/^/multicast /"
