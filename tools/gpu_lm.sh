#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lm}
timeout 300 python tools/lm_attn_check.py check > gpurun_out/${P}_lm_check.log 2>&1
grep -E "LM check|unimp|Error|error" gpurun_out/${P}_lm_check.log | head -20
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or attn or perceiver or vit or lm_" > gpurun_out/${P}_pytest.log 2>&1
tail -3 gpurun_out/${P}_pytest.log
timeout 300 python tools/lm_attn_check.py timeline > gpurun_out/${P}_ff_timeline.log 2>&1
grep -E "^FF|unimp|Error|error" gpurun_out/${P}_ff_timeline.log | head -30
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | grep -v "perceiver_attn_bwd\|eager_perc"
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | grep -v "perceiver_attn_bwd\|eager_perc"
