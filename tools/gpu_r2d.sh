#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2d}
timeout 180 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline.log 2>&1
echo "xb timeline rc=$?"; grep -E "XB|unimp:|Error|error" gpurun_out/${P}_xb_timeline.log | head -40
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_reference_golden.py -m gpu -q -p no:cacheprovider -k "masked_cross or xattn or unmasked or bf16_attention or vit or vision" > gpurun_out/${P}_ktests.log 2>&1
echo "ktests rc=$?"; tail -n 4 gpurun_out/${P}_ktests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_ktests.log | head -20
for wl in C2-rec C3-multitask; do
  timeout 200 python tools/kbench_cli.py --workload $wl --only vit --tag v2 > gpurun_out/${P}_kbv_${wl}_v2.json 2> gpurun_out/${P}_kbv_${wl}_v2.err
  grep "^KB" gpurun_out/${P}_kbv_${wl}_v2.err
  UNIMP_XATTN_FWD_V1=1 timeout 200 python tools/kbench_cli.py --workload $wl --only vit --no-eager --tag v1 > gpurun_out/${P}_kbv_${wl}_v1.json 2> gpurun_out/${P}_kbv_${wl}_v1.err
  grep "^KB" gpurun_out/${P}_kbv_${wl}_v1.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xattn_fwd_tc_kernel -s 4 -c 2 -o gpurun_out/${P}_ncu_xattn_c3 -f python tools/kbench_cli.py --workload C3-multitask --only xattn --no-eager > gpurun_out/${P}_ncu_xattn.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${P}_ncu_xattn_c3.ncu-rep
