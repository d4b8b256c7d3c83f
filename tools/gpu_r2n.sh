#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2n}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unimp_b200/csrc -o /tmp/commit_probe2 tools/probes/commit_probe2.cu && timeout 60 /tmp/commit_probe2 > gpurun_out/${P}_commit_probe2.log 2>&1
echo "probe rc=$?"; cat gpurun_out/${P}_commit_probe2.log
timeout 120 python tools/xblock_check.py check 2>&1 | grep -E "XB check|unimp|Error" | head -8
timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline.log 2>&1
grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline.log | head -4
timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"
