#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lm}
timeout 300 python tools/lm_attn_check.py check > gpurun_out/${P}_lm_check.log 2>&1
grep -E "LM check|unimp|Error|error" gpurun_out/${P}_lm_check.log | head -20
timeout 300 python tools/lm_attn_check.py bench > gpurun_out/${P}_lm_bench.log 2>&1
grep -E "LM bench|unimp|Error|error" gpurun_out/${P}_lm_bench.log | head
