#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2g}
for f in 0 5; do
  UNIMP_XB_FLAGS=$f timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline_f$f.log 2>&1
  echo "flags=$f"; grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline_f$f.log | head -4
  UNIMP_XB_FLAGS=$f timeout 120 python tools/xblock_check.py check 2>&1 | grep -E "XB check" | head -3
  UNIMP_XB_FLAGS=$f timeout 120 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"
done
