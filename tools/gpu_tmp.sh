#!/usr/bin/env bash
mkdir -p gpurun_out
P=r2n2; N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py > gpurun_out/${P}_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -E "OK|Error|error|assert" gpurun_out/${P}_dp_check.log | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-kernel-profile > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err
echo "bench c2 N=$N rc=$?"; tail -c 300 gpurun_out/${P}_bench_c2.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${P}_bench_c2.json") if l.startswith("{")][-1])
    print("C2 N=%d samples/s" % d["n_gpus"], round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
except Exception as e:
    print("no json", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/${P}_ref.json 2> gpurun_out/${P}_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/${P}_ref.json
