// TMA issue-rate probe (dev tool): how fast can threads ISSUE cp.async.bulk.tensor loads?
//   mode 0: one thread, back-to-back loads onto ONE barrier (no waits in the loop)
//   mode 1: W warps, lane 0 of each issues its own stream of loads (own barrier)
//   mode 2: one warp, L lanes each issue (own barrier per lane)
//   mode 3: one thread, mbarrier arrive.expect_tx + try_wait only (no TMA) per iteration
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap tm, long long* out, int n_box,
                                                int box_bytes, int mode, int nthr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int me = -1;
  if (mode == 0 || mode == 3) me = threadIdx.x == 0 ? 0 : -1;
  if (mode == 1) me = (lane == 0 && warp < nthr) ? warp : -1;
  if (mode == 2) me = (warp == 0 && lane < nthr) ? lane : -1;
  long long t0 = clock64();
  if (me >= 0) {
    // all loads of a thread land in the same 2 slots (data is irrelevant), ONE barrier phase per thread
    if (mode != 3) {
      // expect_tx in slices: tx-count is limited to 2^20-1 per phase
      const int per_phase = 16;
      for (int i = 0; i < n_box; i += per_phase) {
        mbar_arrive_expect_tx(&bar[me], per_phase * box_bytes);
        for (int j = 0; j < per_phase; ++j)
          tma_load_2d(smem + (me * 2 + (j & 1)) * box_bytes, &tm, &bar[me], ((i + j) % 40) * 64, (blockIdx.x % 8) * 64);
        if (i == 0) out[148 + blockIdx.x] = clock64() - t0;   // time to ISSUE the first 16
        mbar_wait(&bar[me], (i / per_phase) & 1);
      }
    } else {
      for (int i = 0; i < n_box; ++i) {
        mbar_arrive(&bar[0]);
        mbar_wait(&bar[0], i & 1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int ROWS = 4096, COLS = 2560;
  __nv_bfloat16* d;
  cudaMalloc(&d, (size_t)ROWS * COLS * 2);
  cudaMemset(d, 0, (size_t)ROWS * COLS * 2);
  long long* out;
  cudaMalloc(&out, 2 * 148 * 8);
  for (int box_rows : {16, 64, 256})
    for (int mode : {0, 1, 2, 3})
      for (int nthr : {1, 2, 4, 8}) {
        if ((mode == 0 || mode == 3) && nthr > 1) continue;
        if (mode == 3 && box_rows != 64) continue;
        for (int blocks : {1, 148}) {
          CUtensorMap tm;
          cuuint64_t dims[2] = {COLS, ROWS};
          cuuint64_t strides[1] = {COLS * 2};
          cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
          enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          const int box_bytes = box_rows * 128, n_box = 256;
          const int smem = 1024 + 16 * box_bytes;
          cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          for (int rep = 0; rep < 2; ++rep) probe<<<blocks, 256, smem>>>(tm, out, n_box, box_bytes, mode, nthr);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[296];
          cudaMemcpy(h, out, 296 * 8, cudaMemcpyDeviceToHost);
          double avg = 0, first = 0;
          for (int i = 0; i < blocks; ++i) { avg += h[i]; first += h[148 + i]; }
          avg /= blocks; first /= blocks;
          const int issuers = (mode == 0 || mode == 3) ? 1 : nthr;
          printf("TMA2 mode %d box %3d rows issuers %d SMs %3d: %7.1f cyc per box per issuer, %6.1f B/cyc/SM; first 16 issued in %6.0f cyc %s\n",
                 mode, box_rows, issuers, blocks, avg / n_box, (double)box_bytes * n_box * issuers / avg, first,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  return 0;
}
