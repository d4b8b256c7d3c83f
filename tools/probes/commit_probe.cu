// tcgen05.commit / mbarrier cost probe (dev tool): cycles per loop iteration of the MMA-issuer thread
//   mode 0: commit only                       mode 1: commit + wait on that barrier
//   mode 2: 4 MMAs (N=64) + commit            mode 3: try_wait on an already-completed barrier
//   mode 4: 4 MMAs + commit + wait            mode 5: 4 MMAs only
//   mode 6: 4 MMAs + commit, 6 barriers in rotation, a SECOND thread waits on them and re-arms nothing
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

__global__ void __launch_bounds__(128, 1) probe(long long* out, int n, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8], done;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = make_idesc(128, 64, 0, 0);
  const uint32_t sa = smem_u32(smem), sb = sa + 16384;
  if (threadIdx.x == 0) {
    if (mode == 3) { mbar_arrive(&bar[0]); }
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      if (mode == 2 || mode == 4 || mode == 5 || mode == 6)
        for (int k4 = 0; k4 < 4; ++k4)
          umma_ss(tmem, make_smem_desc(sa + k4 * 32, 16, 1024), make_smem_desc(sb + k4 * 32, 16, 1024), idesc, 1);
      if (mode == 0 || mode == 2) umma_commit(&bar[i & 7]);
      if (mode == 1 || mode == 4) { umma_commit(&bar[0]); mbar_wait(&bar[0], i & 1); }
      if (mode == 3) { if (!mbar_try_wait(&bar[0], 0)) break; }
      if (mode == 6) umma_commit(&bar[i % 6]);
    }
    long long t1 = clock64();
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[0] = t1 - t0;
    out[1] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 1024 + 24576, n = 512;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"commit only", "commit + wait", "4 MMA + commit", "try_wait (complete)", "4 MMA + commit + wait",
                         "4 MMA only", "4 MMA + commit (6 barriers)"};
  for (int mode = 0; mode < 7; ++mode) {
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, smem>>>(d, n, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("COMMIT mode %d %-28s: %7.1f cyc/iter issue, %7.1f cyc/iter incl. drain %s\n", mode, names[mode],
           (double)h[0] / n, (double)h[1] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
