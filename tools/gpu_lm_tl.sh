#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lmt}
timeout 300 python tools/lm_attn_check.py bwd_timeline > gpurun_out/${P}_bwd_timeline.log 2>&1
grep -E "^BW|unimp|Error|error" gpurun_out/${P}_bwd_timeline.log | head -60
