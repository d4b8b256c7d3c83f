#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2j}
for f in 5 13 23 31; do
  UNIMP_XB_FLAGS=$f timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline_f$f.log 2>&1
  echo "flags=$f"; grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline_f$f.log | head -2
done
