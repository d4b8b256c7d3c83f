#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2e}
for f in 0 1 2; do
  UNIMP_XB_FLAGS=$f timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline_f$f.log 2>&1
  echo "flags=$f"; grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline_f$f.log | head -4
done
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_reference_golden.py -m gpu -q -p no:cacheprovider -k "masked_cross or xattn or unmasked or bf16_attention or vit or vision" > gpurun_out/${P}_ktests.log 2>&1
echo "ktests rc=$?"; tail -n 4 gpurun_out/${P}_ktests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_ktests.log | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2_tc_kernel -s 3 -c 1 -o gpurun_out/${P}_ncu_vit -f python tools/kbench_cli.py --workload C2-rec --only vit --no-eager > gpurun_out/${P}_ncu_vit.log 2>&1
echo "ncu vit rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xattn_block_fwd_kernel -s 3 -c 1 -o gpurun_out/${P}_ncu_xb -f python tools/xblock_check.py bench > gpurun_out/${P}_ncu_xb.log 2>&1
echo "ncu xb rc=$?"; ls -la gpurun_out/${P}_ncu_*.ncu-rep
