#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-lms}
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -20
for w in C2-rec C3-multitask; do
for v in 1 0; do
  UNIMP_LM_ATTN=$v timeout 600 python bench.py --workload $w --steps 24 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_${w}_lm$v.json 2> gpurun_out/${P}_bench_${w}_lm$v.err
  echo "bench $w lm=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${P}_bench_${w}_lm$v.json')); print('$w LM_ATTN=$v samples/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), d['clocks'], d.get('gpu_launches'))"
done
done
