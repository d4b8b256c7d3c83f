#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2r}
true
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -20
timeout 400 python bench.py --steps 40 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_nodefer2.json 2> gpurun_out/${P}_bench_nodefer2.err
echo "defer rc=$?"; tail -c 300 gpurun_out/${P}_bench_defer.err
timeout 400 python bench.py --steps 40 --no-cpu-baseline --no-eager-baseline --no-kernel-profile --defer-optimizer > gpurun_out/${P}_bench_nodefer.json 2> gpurun_out/${P}_bench_nodefer.err
echo "nodefer rc=$?"
python - <<PY
import json
for n in ("defer", "nodefer"):
    try:
        d = json.load(open("gpurun_out/${P}_bench_%s.json" % n))
        print(n, "samples/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
    except Exception as e:
        print(n, "no json", e)
PY
true
