#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2k}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unimp_b200/csrc -o /tmp/tma_probe2 tools/probes/tma_probe2.cu -lcuda && timeout 120 /tmp/tma_probe2 > gpurun_out/${P}_tma_probe2.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_tma_probe2.log
bash tools/gpu_r2j.sh ${P}
