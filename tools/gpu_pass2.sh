#!/usr/bin/env bash
# Shorter final pass (dev tool): full GPU suite, smoke, bench lines for configs[1,2,4], launch list of a C3 step, attention kbench.
mkdir -p gpurun_out
P=${1:-r2y}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${P}_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_tests.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/${P}_tests.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/${P}_smoke.log
timeout 700 python bench.py > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err
echo "bench c2 rc=$?"
timeout 700 python bench.py --workload C3-multitask --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
echo "bench c3 rc=$?"
timeout 700 python bench.py --workload C5-imggen --steps 20 --no-cpu-baseline > gpurun_out/${P}_bench_c5.json 2> gpurun_out/${P}_bench_c5.err
echo "bench c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_c3.csv \
  python bench.py --workload C3-multitask --steps 1 --warmup 1 --soak-s 0 --no-cpu-baseline --no-eager-baseline --no-kernel-profile --ncu-range > gpurun_out/${P}_ncu_bench_c3.log 2>&1
echo "ncu launches c3 rc=$?"; wc -l gpurun_out/${P}_launches_c3.csv
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" > gpurun_out/${P}_kbench_attn.log
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm vit --tag $P 2>&1 >/dev/null | grep "^KB" >> gpurun_out/${P}_kbench_attn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'flash_fwd3_kernel' -s 2 -c 2 -o gpurun_out/${P}_ncu_f3_vit -f python tools/kbench_cli.py --workload C3-multitask --only vit --no-eager > gpurun_out/${P}_ncu_f3_vit.log 2>&1
echo "ncu f3 vit rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'flash_fwd3_kernel' -s 2 -c 2 -o gpurun_out/${P}_ncu_f3_lm -f python tools/kbench_cli.py --workload C3-multitask --only lm --no-eager > gpurun_out/${P}_ncu_f3_lm.log 2>&1
echo "ncu f3 lm rc=$?"
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
    for k in ("vit_attn_fwd", "eager_vit_attn_fwd_sdpa", "perceiver_attn_fwd", "lm_attn_fwd", "lm_attn_bwd", "eager_lm_attn_fwd", "eager_lm_attn_bwd", "eager_lm_attn_fwd_causal_only", "eager_lm_attn_bwd_causal_only"):
        v = d["kernels"].get(k)
        if v: print("      ", k, round(v["avg_us"], 2), "us", "x_eager", round(v.get("speedup_vs_eager", 0), 2))
PY
