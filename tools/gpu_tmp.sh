#!/usr/bin/env bash
# one GPU pass: new decode kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "decode or linear_small" > gpurun_out/d1_ktests.log 2>&1; echo "ktests rc=$?"; tail -15 gpurun_out/d1_ktests.log
timeout 900 python -m pytest tests/test_decode.py -x -q -m gpu > gpurun_out/d1_dtests.log 2>&1; echo "dtests rc=$?"; tail -15 gpurun_out/d1_dtests.log
timeout 600 python tools/decode_profile.py 34 > gpurun_out/d1_prof.log 2>&1; echo "prof rc=$?"; head -30 gpurun_out/d1_prof.log
