#!/usr/bin/env bash
# K4 kernels: tests, isolated timings, and A/B inside the real step (dev tool, under gpurun)
mkdir -p gpurun_out
P=${1:-lmab}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -20
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | tee gpurun_out/${P}_kbench_c2.log
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" | tee gpurun_out/${P}_kbench_c3.log
if [ "${2:-}" = "step" ]; then
for v in 1 0; do
  UNIMP_LM_ATTN=$v timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_c2_lm$v.json 2> gpurun_out/${P}_bench_c2_lm$v.err
  echo "bench lm=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${P}_bench_c2_lm$v.json')); print('LM_ATTN=$v samples/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), d['clocks'], d.get('gpu_launches'))"
done
fi
