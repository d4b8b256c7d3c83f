#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/gemv_bench.py 2>&1 | tail -6 > gpurun_out/d6_gemv.log; cat gpurun_out/d6_gemv.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "linear_small or decode" 2>&1 | tail -2
timeout 600 python tools/decode_profile.py 34 > gpurun_out/d6_prof.log 2>&1; echo "prof rc=$?"; sed -n 4,4p gpurun_out/d6_prof.log | cut -c1-150
UNIMP_PDL=0 timeout 600 python tools/decode_profile.py 34 > gpurun_out/d6_prof_nopdl.log 2>&1; echo "prof rc=$?"; sed -n 4,40p gpurun_out/d6_prof_nopdl.log | cut -c1-160
