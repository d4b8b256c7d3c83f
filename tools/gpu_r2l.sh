#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2l}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unimp_b200/csrc -o /tmp/commit_probe tools/probes/commit_probe.cu && timeout 60 /tmp/commit_probe > gpurun_out/${P}_commit_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_commit_probe.log
