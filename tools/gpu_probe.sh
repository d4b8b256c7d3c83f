#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2w}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unimp_b200/csrc -o /tmp/issue_probe tools/probes/issue_probe.cu && timeout 60 /tmp/issue_probe > gpurun_out/${P}_issue_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_issue_probe.log
