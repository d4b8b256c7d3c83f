#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2m}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -30
timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline.log 2>&1
grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline.log | head -4
timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"
for wl in C2-rec C3-multitask; do
  timeout 200 python tools/kbench_cli.py --workload $wl --only xattn vit --tag v2 > gpurun_out/${P}_kb_${wl}.json 2> gpurun_out/${P}_kb_${wl}.err
  grep "^KB" gpurun_out/${P}_kb_${wl}.err
done
