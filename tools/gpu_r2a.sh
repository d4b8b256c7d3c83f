#!/usr/bin/env bash
# Round-2 first GPU pass (dev tool, run under gpurun): tests, the K5 wip harness, and the bench
# lines for configs[1], [2], [4] (train) and [3] (decode).  Everything lands in gpurun_out/r2a_*.
mkdir -p gpurun_out
P=r2a
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider -s > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/${P}_tests.log | head -40
grep -E "4B bf16 vs" gpurun_out/${P}_tests.log
timeout 600 python bench.py > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err
echo "bench c2 rc=$?"; tail -c 400 gpurun_out/${P}_bench_c2.err
timeout 600 python bench.py --workload C3-multitask --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
echo "bench c3 rc=$?"; tail -c 400 gpurun_out/${P}_bench_c3.err
timeout 600 python bench.py --workload C5-imggen --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c5.json 2> gpurun_out/${P}_bench_c5.err
echo "bench c5 rc=$?"; tail -c 400 gpurun_out/${P}_bench_c5.err
true
echo "decode rc=$?"; tail -c 400 gpurun_out/${P}_decode_c4.err
python - <<'PY'
import json
for n in ("c2", "c3", "c5"):
    try:
        d = json.load(open(f"gpurun_out/r2a_bench_{n}.json"))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, "samples/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2),
          d["clocks"], "eager", d.get("gpu_eager_baseline") and d["gpu_eager_baseline"].get("value"),
          "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
    for k, v in d["kernels"].items():
        print("   ", k, round(v["avg_us"], 2), "us", round(v["GB/s"]), "GB/s", round(v["frac_of_hbm_peak"], 3),
              "x_eager", round(v.get("speedup_vs_eager", 0), 2))
try:
    d = json.load(open("gpurun_out/r2a_decode_c4.json"))
    print("decode", d["value"], d["median_ms"], d["end_to_end_tokens_per_s"], d["hf_generate_path"])
except Exception as e:
    print("decode no json", e)
PY
