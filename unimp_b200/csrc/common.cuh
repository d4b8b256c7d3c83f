// Shared device/host helpers for the unimp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/unimp_b200.h"

#ifndef UNIMP_NUM_SMS
#define UNIMP_NUM_SMS 148  // B200: 2 dies x 74 SMs
#endif

namespace unimp {

void set_error(const char* fmt, ...);

#define UNIMP_CHECK_ARG(cond, code, ...)        \
  do {                                          \
    if (!(cond)) {                              \
      ::unimp::set_error(__VA_ARGS__);          \
      return (code);                            \
    }                                           \
  } while (0)

#define UNIMP_CHECK_LAUNCH()                                              \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) {                                             \
      ::unimp::set_error("%s:%d launch failed: %s", __FILE__, __LINE__,   \
                         cudaGetErrorString(e__));                        \
      return (int)e__;                                                    \
    }                                                                     \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of T: 4 floats or 8 bf16.
template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  uint4 raw;
  __device__ __forceinline__ void load(const T* p) { raw = *reinterpret_cast<const uint4*>(p); }
  // streaming load: data touched once, keep it out of L1
  __device__ __forceinline__ void load_stream(const T* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                 : "l"(p));
  }
  __device__ __forceinline__ void store(T* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const;
  __device__ __forceinline__ void pack(const float* f);
};
template <>
__device__ __forceinline__ void Vec16<float>::unpack(float* f) const {
  f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
  f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
}
template <>
__device__ __forceinline__ void Vec16<float>::pack(const float* f) {
  raw.x = __float_as_uint(f[0]); raw.y = __float_as_uint(f[1]);
  raw.z = __float_as_uint(f[2]); raw.w = __float_as_uint(f[3]);
}
template <>
__device__ __forceinline__ void Vec16<__nv_bfloat16>::unpack(float* f) const {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void Vec16<__nv_bfloat16>::pack(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&p);
  }
  raw.x = w[0]; raw.y = w[1]; raw.z = w[2]; raw.w = w[3];
}

// Programmatic dependent launch (decode path: ~500 small dependent kernels per token).  A kernel
// launched through launch_pdl() may START while its predecessor in the stream is still running; it must
// call pdl_wait() before it reads anything an earlier kernel wrote and before its first global write
// (until then it may only touch data that no kernel of the step writes: weights).  Every kernel calls
// pdl_launch_dependents() at its top so that ITS successor may be scheduled early as well.  Both are
// no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // capi.cu: UNIMP_PDL != 0 (default on)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024; `sh` holds >= 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

}  // namespace unimp
