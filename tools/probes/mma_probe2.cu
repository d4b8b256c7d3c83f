// tcgen05.mma rate by operand layout (dev tool): M = 128, K = 16, one elected lane, accumulating into one
// TMEM tile; operands as the attention kernels use them.
//   V 0  A K-major SW128, B K-major SW128, N = 64      (S = Q K^T, panel 0)
//   V 1  A K-major SW128, B MN-major SW128, N = 64     (O += P V, panel 0;  dQ = dS K)
//   V 2  A K-major SW128, B MN-major SW32,  N = 16     (O += P V, panel 1)
//   V 3  A K-major SW32,  B K-major SW32,   N = 64     (S, panel 1: K = 16 of the 80)
//   V 4  A MN-major SW128 (M = 2 x 64), B MN-major SW128 (N = 2 x 64), N = 128   (dK/dV, panel 0)
//   V 5  A MN-major SW128 (M = 2 x 64), B MN-major SW32 (N = 2 x 16),  N = 32    (dK/dV, panel 1)
//   V 6  A K-major SW128, B K-major SW128, N = 128
//   V 7  A K-major SW128, B MN-major SW128 (N = 2 x 64), N = 128
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

template <int V>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t sa = smem_u32(smem), sb = sa + 49152;
  if (threadIdx.x < 32 && elect_one_sync()) {
    constexpr int N = V == 2 ? 16 : (V == 5 ? 32 : (V == 4 || V == 6 || V == 7 ? 128 : 64));
    constexpr uint32_t idesc = make_idesc(128, N, (V == 4 || V == 5) ? 1 : 0, (V == 1 || V == 2 || V == 4 || V == 5 || V == 7) ? 1 : 0);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da, db;
        if (V == 4 || V == 5) da = make_smem_desc(sa + k * 2048, 16384, 1024);       // [P|dS]: 2 chunks of 64, 16 K-rows per step
        else if (V == 3) da = make_smem_desc32(sa, 16, 256);
        else da = make_smem_desc(sa + k * 32, 16, 1024);
        if (V == 0 || V == 6) db = make_smem_desc(sb + k * 32, 16, 1024);
        else if (V == 1) db = make_smem_desc(sb + k * 2048, 1024, 1024);
        else if (V == 2) db = make_smem_desc32(sb + k * 512, 256, 256);
        else if (V == 3) db = make_smem_desc32(sb, 16, 256);
        else if (V == 4 || V == 7) db = make_smem_desc(sb + k * 2048, 16384, 1024);
        else db = make_smem_desc32(sb + k * 512, 4096, 256);
        umma_ss(tmem, da, db, idesc, 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int V>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 98304, n = 512;
  cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<V><<<148, 128, smem>>>(d, n);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("MMA2 V%d %-58s: issue %6.1f cyc/mma, complete %6.1f cyc/mma %s\n", V, name, (double)h[0] / n, (double)h[1] / n,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<0>("A K-major SW128, B K-major SW128, N=64  (S p0)");
  run<1>("A K-major SW128, B MN-major SW128, N=64 (PV p0, dQ)");
  run<2>("A K-major SW128, B MN-major SW32, N=16  (PV p1)");
  run<3>("A K-major SW32, B K-major SW32, N=64    (S p1)");
  run<4>("A MN-major 2x64, B MN-major 2x64, N=128 (dKV p0)");
  run<5>("A MN-major 2x64, B MN-major SW32 2x16, N=32 (dKV p1)");
  run<6>("A K-major SW128, B K-major SW128, N=128");
  run<7>("A K-major SW128, B MN-major 2x64, N=128");
  return 0;
}
