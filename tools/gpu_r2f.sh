#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2f}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unimp_b200/csrc -o /tmp/mma_probe tools/probes/mma_probe.cu && timeout 120 /tmp/mma_probe > gpurun_out/${P}_mma_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_mma_probe.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -30
