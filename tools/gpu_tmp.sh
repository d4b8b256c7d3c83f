#!/usr/bin/env bash
# final artefacts of round 2 (second session): GPU suite, smoke, bench lines, decode line, launch list
mkdir -p gpurun_out
P=r2z
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/${P}_smoke.log
timeout 700 python bench.py > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err
echo "bench c2 rc=$?"
timeout 700 python bench.py --workload C3-multitask --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
echo "bench c3 rc=$?"
timeout 700 python bench.py --workload C5-imggen --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c5.json 2> gpurun_out/${P}_bench_c5.err
echo "bench c5 rc=$?"
timeout 500 python bench.py --mode decode --steps 5 > gpurun_out/${P}_decode_c4.json 2> gpurun_out/${P}_decode_c4.err
echo "decode rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches.csv \
  python bench.py --steps 1 --warmup 1 --soak-s 0 --no-cpu-baseline --no-eager-baseline --no-kernel-profile --ncu-range > gpurun_out/${P}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/${P}_launches.csv
python - <<PY
import json
for n in ("c2", "c3", "c5"):
    try:
        d = json.load(open("gpurun_out/${P}_bench_%s.json" % n))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, "samples/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2),
          d["clocks"], "eager", d.get("gpu_eager_baseline") and d["gpu_eager_baseline"].get("value"),
          "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
    for k in ("roofline", "roofline_core", "roofline_dominant"):
        r = d.get(k)
        if r: print("   ", k, r["kernel"][:40], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "avg_us", round(r["avg_us"], 2))
try:
    d = json.load(open("gpurun_out/${P}_decode_c4.json"))
    print("decode", d["value"], d["ms_per_token"], d["roofline"]["frac"], d["median_ms"], d["hf_generate_path"])
    for k, v in d["kernels"].items():
        print("      ", k, round(v["avg_us"], 2), "us", round(v["GB/s"]), "GB/s", round(v["frac_of_hbm_peak"], 3), "x_eager", round(v.get("speedup_vs_eager", 0), 2))
except Exception as e:
    print("decode no json", e)
PY
