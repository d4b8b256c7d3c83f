#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-r2r}
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "deferred or graphed or adamw or unmasked or vision or attention" > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -20
timeout 400 python bench.py --steps 40 --no-cpu-baseline --no-eager-baseline --no-kernel-profile > gpurun_out/${P}_bench_defer.json 2> gpurun_out/${P}_bench_defer.err
echo "defer rc=$?"; tail -c 300 gpurun_out/${P}_bench_defer.err
timeout 400 python bench.py --steps 40 --no-cpu-baseline --no-eager-baseline --no-kernel-profile --no-defer-optimizer > gpurun_out/${P}_bench_nodefer.json 2> gpurun_out/${P}_bench_nodefer.err
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
timeout 200 python tools/kbench_cli.py --workload C2-rec --only misc --no-eager --tag bg 2>&1 | grep "^KB"
