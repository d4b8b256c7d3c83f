#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lmt}
timeout 300 python tools/lm_attn_check.py timeline > gpurun_out/${P}_ff_timeline.log 2>&1
grep -E "^FF|unimp|Error|error" gpurun_out/${P}_ff_timeline.log | head -30
