#!/usr/bin/env bash
# Build and run one hardware micro-probe under gpurun: gpu_probe.sh <tag> <probe name without .cu> [extra nvcc flags]
mkdir -p gpurun_out
P=${1:-probe}; N=${2:-mma_probe2}; shift; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 "$@" -I unimp_b200/csrc -o /tmp/$N tools/probes/$N.cu && timeout 60 /tmp/$N > gpurun_out/${P}_$N.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_$N.log
