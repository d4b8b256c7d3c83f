#!/usr/bin/env bash
# compute-sanitizer over the K2/K3/K4 kernels on small shapes (dev tool, under gpurun)
mkdir -p gpurun_out
P=${1:-san}
for f3 in 0 1; do
for tool in memcheck racecheck; do
  UNIMP_FLASH3=$f3 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/lm_attn_check.py small > gpurun_out/${P}_${tool}_f3_$f3.log 2>&1
  echo "flash3=$f3 $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LM check|SMALL|Error|hazard" gpurun_out/${P}_${tool}_f3_$f3.log | head -12
done
done
