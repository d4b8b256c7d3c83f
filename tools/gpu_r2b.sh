#!/usr/bin/env bash
# Round-2 kernel pass (dev tool, under gpurun): K1 v2 + K1-fused correctness and isolated timings.
mkdir -p gpurun_out
P=${1:-r2b}
timeout 180 python tools/xblock_check.py check > gpurun_out/${P}_xb_check.log 2>&1
echo "xb check rc=$?"; grep -E "XB|unimp:|Error|error" gpurun_out/${P}_xb_check.log | head -30
timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "masked_cross or xattn or unmasked or bf16_attention" > gpurun_out/${P}_ktests.log 2>&1
echo "ktests rc=$?"; tail -n 4 gpurun_out/${P}_ktests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_ktests.log | head -20
for wl in C2-rec C3-multitask; do
  timeout 200 python tools/kbench_cli.py --workload $wl --only xattn --no-eager --tag v2 > gpurun_out/${P}_kb_${wl}_v2.json 2> gpurun_out/${P}_kb_${wl}_v2.err
  grep "^KB" gpurun_out/${P}_kb_${wl}_v2.err
  UNIMP_XATTN_FWD_V1=1 timeout 200 python tools/kbench_cli.py --workload $wl --only xattn --no-eager --tag v1 > gpurun_out/${P}_kb_${wl}_v1.json 2> gpurun_out/${P}_kb_${wl}_v1.err
  grep "^KB" gpurun_out/${P}_kb_${wl}_v1.err
done
timeout 300 python tools/xblock_check.py bench > gpurun_out/${P}_xb_bench.log 2>&1
echo "xb bench rc=$?"; grep -E "XB|unimp:|Error" gpurun_out/${P}_xb_bench.log | head
