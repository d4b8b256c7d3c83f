#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2p}
timeout 120 python tools/xblock_check.py check 2>&1 | grep -E "XB check|unimp|Error" | head -8
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "xattn_block" > gpurun_out/${P}_ktests.log 2>&1
echo "ktests rc=$?"; tail -n 3 gpurun_out/${P}_ktests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_ktests.log | head
timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline.log 2>&1
grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline.log | head -4
timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"
