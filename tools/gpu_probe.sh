#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2w}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unimp_b200/csrc -o /tmp/mma_probe2 tools/probes/mma_probe2.cu && timeout 60 /tmp/mma_probe2 > gpurun_out/${P}_mma_probe2.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_mma_probe2.log
