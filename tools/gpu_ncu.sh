#!/usr/bin/env bash
# ncu --set full captures of the attention kernels at configs[1] / configs[2] shapes (dev tool, under gpurun).
mkdir -p gpurun_out
P=${1:-r2u}
cap() {  # name, kernel regex, skip, count, workload, sections
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/${P}_ncu_$1 -f \
    python tools/kbench_cli.py --workload $5 --only $6 --no-eager > gpurun_out/${P}_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap xblock_c2 xattn_block_fwd_kernel 4 2 C2-rec xattn
cap xblock_c3 xattn_block_fwd_kernel 4 2 C3-multitask xattn
cap xcore_c2 xattn_fwd_tc_kernel 4 2 C2-rec xattn
cap xbwd_c2 attn_bwd_tc_kernel 2 2 C2-rec xattn
cap vit_c2 flash_fwd_kernel 4 2 C2-rec vit
cap k5_c2 gate_residual_ln 4 8 C2-rec k5
ls -la gpurun_out/${P}_ncu_*.ncu-rep
