#!/usr/bin/env bash
# One GPU-box pass (dev tool, run under gpurun): GPU tests, isolated K5 / rotary timings, the bench
# line, and (NCU=1) one `ncu --set full` capture of the K5 / GELU kernels.
# Everything lands in gpurun_out/ with the prefix given as $1.
P=${1:-x}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/${P}_tests.log | head -20
timeout 90 python tools/ln_bench.py > gpurun_out/${P}_lnbench.log 2>&1
grep LNBENCH gpurun_out/${P}_lnbench.log || tail -n 5 gpurun_out/${P}_lnbench.log
timeout 280 python bench.py ${BENCH_ARGS:-} > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${P}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${P}_bench.json"))
print("samples/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], "cpu", d.get("cpu_baseline"))
for k, v in d["kernels"].items():
    print(" ", k, round(v["avg_us"], 2), "us", round(v["GB/s"]), "GB/s", round(v["frac_of_hbm_peak"], 3))
PY
if [ "${NCU:-0}" = "1" ]; then
  B=6 REPS=1 timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:'gate_residual_ln|gelu_' -c 24 -o gpurun_out/${P}_prof -f python tools/run_kernels.py > gpurun_out/${P}_ncu.log 2>&1
  echo "ncu rc=$?"; ls -la gpurun_out/${P}_prof.ncu-rep
fi
