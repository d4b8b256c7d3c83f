#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "beam_topk" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_decode.py tests/test_model_gpu.py -x -q -m gpu -k "decode or graphed or generate" 2>&1 | tail -3
timeout 600 python bench.py --mode decode --no-kernel-profile > gpurun_out/h1_decode.json 2> gpurun_out/h1_decode.err; echo "bench rc=$?"
UNIMP_BEAM_TOPK=0 timeout 600 python bench.py --mode decode --no-kernel-profile > gpurun_out/h1_decode_torch_topk.json 2> gpurun_out/h1_decode_torch_topk.err; echo "bench rc=$?"
python - <<'P'
import json
for f in ('h1_decode','h1_decode_torch_topk'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['ms_per_token'],4), round(d['roofline']['frac'],3))
P
