#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode.py tests/test_kernels_gpu.py -x -q -m gpu -k "decode or linear_small or graphed" 2>&1 | tail -2
timeout 900 python bench.py --mode decode > gpurun_out/d9_decode.json 2> gpurun_out/d9_decode.err; echo "bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/d9_decode.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_token'], d['roofline'], d['median_ms'], d['hf_generate_path'])
P
tail -3 gpurun_out/d9_decode.err
UNIMP_PDL=0 timeout 600 python tools/decode_profile.py 34 > gpurun_out/d9_prof_nopdl.log 2>&1; sed -n 1,1p gpurun_out/d9_prof_nopdl.log;  sed -n 4,12p gpurun_out/d9_prof_nopdl.log | cut -c1-130
