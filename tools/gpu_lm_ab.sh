#!/usr/bin/env bash
# A/B of the K4 kernels inside the real step (dev tool, under gpurun)
mkdir -p gpurun_out
P=${1:-lmab}
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -20
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm --tag $P 2>&1 >/dev/null | grep "^KB"
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm --tag $P 2>&1 >/dev/null | grep "^KB"
for v in 1 0; do
  UNIMP_LM_ATTN=$v timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_c2_lm$v.json 2> gpurun_out/${P}_bench_c2_lm$v.err
  echo "bench lm=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${P}_bench_c2_lm$v.json')); print('LM_ATTN=$v samples/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), d['clocks'], d.get('gpu_launches'))"
done
for v in 1 0; do
  UNIMP_LM_ATTN=$v timeout 600 python bench.py --workload C3-multitask --steps 12 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_c3_lm$v.json 2> gpurun_out/${P}_bench_c3_lm$v.err
  echo "bench c3 lm=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${P}_bench_c3_lm$v.json')); print('C3 LM_ATTN=$v samples/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), d['clocks'])"
done
