#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2h}
for f in 0 5; do
  UNIMP_XB_FLAGS=$f timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline_f$f.log 2>&1
  echo "flags=$f"; grep -E "per chunk|XB   0:|unimp:|Error" gpurun_out/${P}_xb_timeline_f$f.log | head -4
done
