#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode.py tests/test_kernels_gpu.py -x -q -m gpu -k "decode or graphed" 2>&1 | tail -3
timeout 300 python tools/kbench_cli.py --only decode --no-eager 2>&1 | grep "^KB" | grep "attn"
UNIMP_DECODE_ATTN_SPLIT=1 timeout 300 python tools/kbench_cli.py --only decode --no-eager 2>&1 | grep "^KB" | grep "lm_decode"
timeout 900 python bench.py --mode decode --no-kernel-profile > gpurun_out/e2_decode.json 2> gpurun_out/e2_decode.err; echo "bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/e2_decode.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_token'], d['roofline']['frac'], d['median_ms'])
P
