#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lmt}
timeout 300 python tools/lm_attn_check.py timeline3 > gpurun_out/${P}_f3_timeline.log 2>&1
grep -E "^F3|unimp|Error|error" gpurun_out/${P}_f3_timeline.log | head -40
