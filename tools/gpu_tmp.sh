mkdir -p gpurun_out
P=r2g
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${P}_smoke.log
timeout 700 python bench.py --workload C3-multitask --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
echo "bench c3 rc=$?"
timeout 700 python bench.py --workload C5-imggen --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c5.json 2> gpurun_out/${P}_bench_c5.err
echo "bench c5 rc=$?"
timeout 700 python bench.py > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err
echo "bench c2 rc=$?"
python - <<PY
import json
for n in ("c2", "c3", "c5"):
    try:
        d = json.load(open("gpurun_out/r2g_bench_%s.json" % n))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, "samples/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2),
          d["clocks"], "eager", d.get("gpu_eager_baseline") and d["gpu_eager_baseline"].get("value"),
          "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
    r = d["roofline"]; print("    roofline", r["kernel"][:44], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
    print("    fused launches/step", d["gpu_launches_per_step"].get("unimp_xattn_block_fwd"), "core", d["gpu_launches_per_step"].get("unimp_xattn_fwd"))
PY
