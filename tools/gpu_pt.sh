#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-pt}
timeout 300 python tools/lm_attn_check.py check 2>&1 | grep -E "LM check|unimp|rror" | tail -7
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or attn or perceiver or vit or lm_" 2>&1 | tail -2
UNIMP_FLASH3=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or attn or perceiver or vit or lm_" 2>&1 | tail -2
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | grep "fwd" | tee gpurun_out/${P}_kbench.log
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | grep "fwd" | tee -a gpurun_out/${P}_kbench.log
echo "UNIMP_FLASH_PT=0 (flash3, C3)"
UNIMP_FLASH_PT=0 timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P --no-eager 2>&1 >/dev/null | grep "^KB" | grep "fwd"
