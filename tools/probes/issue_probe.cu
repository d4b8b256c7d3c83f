// Does the MMA issuer overlap its mbarrier wait with tensor-pipe execution?  (dev tool)
// One elected lane, fully unrolled bodies (constant descriptors), per iteration:
//   mode 0: 8 MMAs (N=64)          mode 1: 8 MMAs + commit        mode 2: 8 MMAs + commit + try_wait(complete)
//   mode 3: try_wait(complete) only mode 4: 16 MMAs + commit + try_wait   mode 5: 8 MMAs N=128 + commit + try_wait
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8], done, ready;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    mbar_init(&done, 1); mbar_init(&ready, 1);
    fence_barrier_init();
    mbar_arrive(&ready);
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 128);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  constexpr int N = MODE == 5 ? 128 : 64;
  constexpr int NM = MODE == 4 ? 16 : (MODE == 3 ? 0 : 8);
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  const uint32_t sa = smem_u32(smem), sb = sa + 16384;
  if (threadIdx.x >= 32 && threadIdx.x < 64 && elect_one_sync()) {
    long long t0 = clock64();
    for (int i = 0; i < n; i += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int k = 0; k < NM; ++k)
          umma_ss(tmem, make_smem_desc(sa + (k & 3) * 32, 16, 1024), make_smem_desc(sb + (k & 3) * 32, 16, 1024), idesc, 1);
        if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5) umma_commit(&bar[u]);
        if (MODE == 2 || MODE == 3 || MODE == 4 || MODE == 5) { if (!mbar_try_wait(&ready, 0)) break; }
      }
    }
    long long t1 = clock64();
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[0] = t1 - t0;
    out[1] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 128);
}

template <int MODE>
void run(const char* name, int floor_cyc) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 1024 + 49152, n = 512;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<MODE><<<1, 128, smem>>>(d, n);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("ISSUE mode %d %-36s: %7.1f cyc/iter issue side, %7.1f cyc/iter incl. drain (tensor floor %d) %s\n", MODE, name,
         (double)h[0] / n, (double)h[1] / n, floor_cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<0>("8 MMA N=64", 384);
  run<1>("8 MMA N=64 + commit", 384);
  run<2>("8 MMA N=64 + commit + try_wait", 384);
  run<3>("try_wait(complete) only", 0);
  run<4>("16 MMA N=64 + commit + try_wait", 768);
  run<5>("8 MMA N=128 + commit + try_wait", 512);
  return 0;
}
