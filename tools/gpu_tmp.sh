#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/xblock_check.py check > gpurun_out/x1_check.log 2>&1; echo "check rc=$?"; grep "XB check" gpurun_out/x1_check.log; tail -3 gpurun_out/x1_check.log | grep -v "XB check"
timeout 300 python tools/xblock_check.py timeline > gpurun_out/x1_timeline.log 2>&1; grep "XB" gpurun_out/x1_timeline.log | head -14
timeout 300 python tools/xblock_check.py bench > gpurun_out/x1_bench.log 2>&1; grep "XB bench" gpurun_out/x1_bench.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "xattn_block or fused" 2>&1 | tail -2
