// What bounds a softmax step on one SM?  (dev tool)  One CTA per SM, W softmax warps, R iterations of
//   mode 0: 2 x tcgen05.ld.32x32b.x32 (+ wait)            -> TMEM read port
//   mode 1: 64 x ex2.approx per thread on register values  -> MUFU
//   mode 2: both, as in flash_fwd (load 64 columns, then 64 exponentials of them)
//   mode 3: mode 2 + pack to bf16 + 8 x 16-byte shared stores + fence.proxy.async (the P store)
// Reported: SM cycles per (warp-quad = 128 rows x 64 columns) block, i.e. per "tile-step".
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(long long* out, float* sink, int R) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  float f0[32], f1[32];
  uint32_t* s0 = reinterpret_cast<uint32_t*>(f0);
  uint32_t* s1 = reinterpret_cast<uint32_t*>(f1);
#pragma unroll
  for (int c = 0; c < 32; ++c) { f0[c] = -(float)(c + threadIdx.x % 7); f1[c] = -(float)(c * 2 + 1); }
  float acc = 0.f;
  uint8_t* sP = smem + (warp >> 2) * 16384;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < R; ++it) {
    if (MODE == 0 || MODE >= 2) {
      tmem_ld32(lane_addr, s0);
      tmem_ld32(lane_addr + 32, s1);
      tmem_ld_wait();
    }
    if (MODE >= 1) {
      uint4 pk[8];
      uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float x0 = MODE == 1 ? f0[c] + acc : f0[c], x1 = MODE == 1 ? f0[c + 1] + acc : f0[c + 1];
        const float y0 = MODE == 1 ? f1[c] + acc : f1[c], y1 = MODE == 1 ? f1[c + 1] + acc : f1[c + 1];
        const float e0 = exp2f(fmaf(x0, 0.16f, -1.f)), e1 = exp2f(fmaf(x1, 0.16f, -1.f));
        const float g0 = exp2f(fmaf(y0, 0.16f, -1.f)), g1 = exp2f(fmaf(y1, 0.16f, -1.f));
        ps[(c >> 1) & 3] += (e0 + e1) + (g0 + g1);
        __nv_bfloat162 a = __floats2bfloat162_rn(e0, e1), b = __floats2bfloat162_rn(g0, g1);
        pw[c >> 1] = *reinterpret_cast<uint32_t*>(&a);
        pw[16 + (c >> 1)] = *reinterpret_cast<uint32_t*>(&b);
      }
      acc = acc * 0.5f + ((ps[0] + ps[1]) + (ps[2] + ps[3])) * 1e-3f;
      if (MODE == 3) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(sP + sw128_offset(threadIdx.x & 127, c)) = pk[c];
        fence_proxy_async_smem();
      } else {
        acc += __uint_as_float(pw[0] ^ pw[17]) * 1e-30f;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + f0[3] + f1[5];
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE>
void run(const char* name, int warps) {
  long long* d;
  float* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 148 * 512 * 4);
  const int R = 256, smem = 65536;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<MODE><<<148, warps * 32, smem>>>(d, sink, R);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += h[i];
  const double per_iter = s / 148 / R;                 // cycles per iteration of the whole CTA
  printf("SOFTMAX mode %d %-34s %2d warps: %7.1f cycles per iteration = %7.1f per 128x64 block per SM\n", MODE, name,
         warps, per_iter, per_iter / (warps / 4));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 12, 16}) run<0>("2 x tcgen05.ld x32 + wait", w);
  for (int w : {4, 8, 12, 16}) run<1>("64 ex2 per thread (+ pack)", w);
  for (int w : {4, 8, 12, 16}) run<2>("ld + 64 ex2 + pack", w);
  for (int w : {4, 8, 12, 16}) run<3>("ld + 64 ex2 + pack + STS + fence", w);
  return 0;
}
