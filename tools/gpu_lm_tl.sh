#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lmt}
timeout 300 python tools/lm_attn_check.py check > gpurun_out/${P}_lm_check.log 2>&1
grep -E "LM check|unimp|Error|error" gpurun_out/${P}_lm_check.log | head -20
timeout 300 python tools/lm_attn_check.py bwd_timeline > gpurun_out/${P}_bwd_timeline.log 2>&1
grep -E "^BW|unimp|Error|error" gpurun_out/${P}_bwd_timeline.log | head -40
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm --tag $P --no-eager 2>&1 >/dev/null | grep "^KB"
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm --tag $P --no-eager 2>&1 >/dev/null | grep "^KB"
