#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-f3}
for sg in 0 400 700 1000; do
echo "UNIMP_FLASH3_STAGGER=$sg"
UNIMP_FLASH3_STAGGER=$sg timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P --no-eager 2>&1 >/dev/null | grep "^KB" | grep "vit_attn_fwd\|lm_attn_fwd"
done
UNIMP_FLASH3_STAGGER=700 timeout 300 python tools/lm_attn_check.py timeline3 > gpurun_out/${P}_f3_timeline.log 2>&1
grep -E "^F3" gpurun_out/${P}_f3_timeline.log | tail -14
