#!/usr/bin/env bash
# 8-GPU pass (dev tool, under gpurun --gpus 8): data-parallel equality check + configs[2] at N=8.
mkdir -p gpurun_out
P=${1:-r2n8}
N=${2:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py > gpurun_out/${P}_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -E "OK|Error|error|assert" gpurun_out/${P}_dp_check.log | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload C3-multitask --steps 20 --warmup 5 > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
echo "bench c3 N=$N rc=$?"; tail -c 300 gpurun_out/${P}_bench_c3.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${P}_bench_c3.json") if l.startswith("{")][-1])
    print("C3 N=%d samples/s" % d["n_gpus"], round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
except Exception as e:
    print("no json", e)
PY
