#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "linear_small" 2>&1 | tail -2
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python tools/decode_kernels_once.py > gpurun_out/san2_racecheck_decode.log 2>&1; echo "racecheck rc=$?"; grep -v "Host Frame" gpurun_out/san2_racecheck_decode.log | tail -12
timeout 300 python tools/kbench_cli.py --only decode --no-eager 2>&1 | grep "^KB" | grep "linear"
