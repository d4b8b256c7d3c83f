// MMA-issuer loop cost probe, elect.sync version (dev tool).  One warp, elected lane, per iteration:
//   mode 0: 4 MMAs (N=64)                       mode 1: 4 MMAs + commit (6 barriers in rotation)
//   mode 2: try_wait on a completed barrier     mode 3: test_wait on a completed barrier
//   mode 4: 4 MMAs + commit + try_wait(completed barrier)      (the real loop's shape)
//   mode 5: 8 MMAs + commit + try_wait(completed)
//   mode 6: raw ld.shared of the barrier word   mode 7: 4 MMAs + commit + raw peek
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128, 1) probe(long long* out, int n, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8], done, ready;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    mbar_init(&done, 1); mbar_init(&ready, 1);
    fence_barrier_init();
    mbar_arrive(&ready);   // phase 0 of `ready` is complete from now on
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = make_idesc(128, 64, 0, 0);
  const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024), db = make_smem_desc(smem_u32(smem) + 16384, 16, 1024);
  if (threadIdx.x >= 32 && threadIdx.x < 64 && elect_one_sync()) {
    long long t0 = clock64();
    unsigned long long sink = 0;
    for (int i = 0; i < n; ++i) {
      const int nm = (mode == 5) ? 8 : ((mode == 0 || mode == 1 || mode == 4 || mode == 7) ? 4 : 0);
      for (int k = 0; k < nm; ++k) umma_ss(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
      if (mode == 1 || mode == 4 || mode == 5 || mode == 7) umma_commit(&bar[i % 6]);
      if (mode == 2 || mode == 4 || mode == 5) { if (!mbar_try_wait(&ready, 0)) break; }
      if (mode == 3) { if (!mbar_test_wait(&ready, 0)) break; }
      if (mode == 6 || mode == 7) sink += *reinterpret_cast<volatile unsigned long long*>(&ready);
    }
    long long t1 = clock64();
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[0] = t1 - t0;
    out[1] = clock64() - t0;
    out[2] = (long long)sink;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  long long* d;
  cudaMalloc(&d, 24);
  const int smem = 1024 + 24576, n = 512;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"4 MMA", "4 MMA + commit", "try_wait (complete)", "test_wait (complete)",
                         "4 MMA + commit + try_wait", "8 MMA + commit + try_wait", "raw ld.shared peek", "4 MMA + commit + peek"};
  for (int mode = 0; mode < 8; ++mode) {
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, smem>>>(d, n, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3];
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("COMMIT2 mode %d %-28s: %7.1f cyc/iter issue, %7.1f cyc/iter incl. drain %s\n", mode, names[mode],
           (double)h[0] / n, (double)h[1] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
