// tcgen05.mma issue-rate probe (dev tool): cycles per UTCHMMA for SS operands in shared memory,
// M = 128, K = 16, various N, with the accumulator chain dependent (same TMEM tile) or spread over
// several tiles.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I unimp_b200/csrc -o /tmp/mma_probe tools/probes/mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

template <int N, int NACC, int KSTEPS_PER_TILE>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 128 x 64 bf16 (16 KB)
  uint8_t* sB = smem + 16384;          // N x 64 bf16
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, 0);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int acc = (i / KSTEPS_PER_TILE) % NACC, k4 = i % 4;
      umma_ss(tmem + acc * N, make_smem_desc(smem_u32(sA) + k4 * 32, 16, 1024),
              make_smem_desc(smem_u32(sB) + k4 * 32, 16, 1024), idesc, 1);
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

template <int N, int NACC, int KS>
void run(const char* name, int n_mma, int blocks) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 1024 + 16384 + N * 128;
  cudaFuncSetAttribute(probe<N, NACC, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<N, NACC, KS><<<blocks, 128, smem>>>(d, n_mma);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("MMA %-34s blocks=%3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (floor %d)%s\n", name, blocks,
         (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int n = 512;
  for (int blocks : {1, 148}) {
    run<64, 1, 1>("N=64  one accumulator", n, blocks);
    run<64, 4, 1>("N=64  4 accumulators rotating", n, blocks);
    run<64, 4, 4>("N=64  4 acc, 4 k-steps each", n, blocks);
    run<128, 1, 1>("N=128 one accumulator", n, blocks);
    run<128, 2, 1>("N=128 2 accumulators rotating", n, blocks);
    run<160, 1, 1>("N=160 one accumulator", n, blocks);
    run<160, 2, 1>("N=160 2 accumulators rotating", n, blocks);
    run<256, 1, 1>("N=256 one accumulator", n, blocks);
    run<256, 2, 1>("N=256 2 accumulators rotating", n, blocks);
  }
  return 0;
}
