#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2o}
timeout 120 python tools/xblock_check.py check 2>&1 | grep -E "XB check|unimp|Error" | head -8
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "xattn_block" > gpurun_out/${P}_ktests.log 2>&1
echo "ktests rc=$?"; tail -n 3 gpurun_out/${P}_ktests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_ktests.log | head
timeout 120 python tools/xblock_check.py timeline > gpurun_out/${P}_xb_timeline.log 2>&1
grep -E "per-CTA durations|unimp:|Error" gpurun_out/${P}_xb_timeline.log | head -4
timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"
UNIMP_XB_FLAGS=1 timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench" | sed 's/^/nomc /'
# step-level A/B: fused vs three-launch x-attn in the real training step
UNIMP_XATTN_FUSED=1 timeout 400 python bench.py --steps 30 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_fused.json 2> gpurun_out/${P}_bench_fused.err
UNIMP_XATTN_FUSED=0 timeout 400 python bench.py --steps 30 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_unfused.json 2> gpurun_out/${P}_bench_unfused.err
python - <<'PY'
import json
for n in ("fused", "unfused"):
    try:
        d = json.load(open(f"gpurun_out/r2o_bench_{n}.json"))
        print(n, "samples/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
    except Exception as e:
        print(n, "no json", e)
PY
