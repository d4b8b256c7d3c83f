#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/kbench_cli.py --only k5 --no-eager 2>&1 | grep "^KB" | grep "ln"
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "gate_residual" 2>&1 | tail -2
UNIMP_PDL=0 timeout 600 python tools/decode_profile.py 34 > gpurun_out/f1_prof_nopdl.log 2>&1; sed -n 1,1p gpurun_out/f1_prof_nopdl.log;  sed -n 4,9p gpurun_out/f1_prof_nopdl.log | cut -c1-130
