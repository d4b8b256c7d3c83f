#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "linear_small" 2>&1 | tail -2
timeout 600 python bench.py --mode decode --no-kernel-profile > gpurun_out/g1_decode.json 2> gpurun_out/g1_decode.err; echo "bench rc=$?"
UNIMP_GEMV_PREFETCH=0 timeout 600 python bench.py --mode decode --no-kernel-profile > gpurun_out/g1_decode_nopf.json 2> gpurun_out/g1_decode_nopf.err; echo "bench rc=$?"
python - <<'P'
import json
for f in ('g1_decode','g1_decode_nopf'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['ms_per_token'],4), round(d['roofline']['frac'],3))
P
