#!/usr/bin/env bash
# Builds unimp_b200/libunimp_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libunimp_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC -Xcompiler -fvisibility=default ${UNIMP_NVCC_EXTRA:-}"
mkdir -p ../../build
objs=()
pids=()
for f in capi focal_ce gate_ln misc attn_simt attn_tc xattn_block lm_attn flash_fwd lm_fused decode; do
  o=../../build/$f.o
  if [ ! -f "$o" ] || [ "$f.cu" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ ../../include/unimp_b200.h -nt "$o" ] || { [ -f tc_common.cuh ] && [ tc_common.cuh -nt "$o" ]; }; then
    $NVCC $FLAGS -c $f.cu -o $o &
    pids+=($!)
  fi
  objs+=($o)
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT "${objs[@]}" -lcudart
echo "built $OUT"
