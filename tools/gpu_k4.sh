#!/usr/bin/env bash
# K2/K3/K4 attention kernels, one-call dev loop under gpurun: parity checks, the attention tests, step
# timelines (clock64 stamps), isolated timings with both eager columns, and the A/B switches:
#   UNIMP_FLASH3=0|1 (two-CTA vs persistent three-tile forward)   UNIMP_FLASH_PT=0 (P through shared memory)
#   UNIMP_LM_BWD_FLAGS=1|3|4 (dQ path / item order experiments)   UNIMP_LM_ATTN=0 (cuDNN SDPA in the model)
mkdir -p gpurun_out
P=${1:-k4}
timeout 300 python tools/lm_attn_check.py check > gpurun_out/${P}_lm_check.log 2>&1
grep -E "LM check|unimp|Error|error" gpurun_out/${P}_lm_check.log | head -20
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or attn or perceiver or vit or lm_" > gpurun_out/${P}_pytest.log 2>&1
tail -3 gpurun_out/${P}_pytest.log
UNIMP_FLASH3=0 timeout 300 python tools/lm_attn_check.py timeline > gpurun_out/${P}_ff_timeline.log 2>&1
UNIMP_FLASH3=1 timeout 300 python tools/lm_attn_check.py timeline3 > gpurun_out/${P}_f3_timeline.log 2>&1
timeout 300 python tools/lm_attn_check.py bwd_timeline > gpurun_out/${P}_bwd_timeline.log 2>&1
grep -hE "^FF timeline|^F3 timeline|^BW timeline|unimp:|Error" gpurun_out/${P}_ff_timeline.log gpurun_out/${P}_f3_timeline.log gpurun_out/${P}_bwd_timeline.log | head
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | tee gpurun_out/${P}_kbench.log
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | tee -a gpurun_out/${P}_kbench.log
