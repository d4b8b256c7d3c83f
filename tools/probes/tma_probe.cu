// TMA tiled-load service-rate probe (dev tool): one thread per CTA streams boxes of a bf16 matrix
// (64 columns x R rows, SWIZZLE_128B) into a ring of shared-memory stages; prints bytes/cycle/SM
// and ns per box for different row strides, box heights, ring depths and numbers of active SMs.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace unimp::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, long long* out, int n_box,
                                               int stages, int box_bytes, int n_col_boxes, int n_row_boxes, int box_rows,
                                               int same_tile) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    const int rb = same_tile ? 0 : (blockIdx.x % n_row_boxes);
    for (int i = 0; i < n_box + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&full[s], ((i / stages) - 1) & 1);   // box i - stages has landed: reuse its slot
      if (i < n_box) {
        mbar_arrive_expect_tx(&full[s], box_bytes);
        tma_load_2d(smem + s * box_bytes, &tm, &full[s], (i % n_col_boxes) * 64, rb * box_rows);
      }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
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
  cudaMalloc(&out, 148 * 8);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int strided = 1; strided >= 0; --strided)
    for (int box_rows : {64, 128, 256})
      for (int stages : {2, 6})
        for (int blocks : {1, 96, 148})
          for (int same : {0, 1}) {
            if (same && blocks == 1) continue;
            CUtensorMap tm;
            // strided: a (ROWS x 2560) matrix, 64-column boxes (row stride 5120 B);
            // contiguous: the same memory viewed as (ROWS*40 x 64): rows 128 B apart
            cuuint64_t dims[2] = {(cuuint64_t)(strided ? COLS : 64), (cuuint64_t)(strided ? ROWS : ROWS * 40)};
            cuuint64_t strides[1] = {(cuuint64_t)(strided ? COLS * 2 : 128)};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int box_bytes = box_rows * 128, n_box = 400;
            const int smem = 1024 + stages * box_bytes;
            cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            const int n_col = strided ? 40 : 1, n_rowb = (strided ? ROWS : ROWS * 40) / box_rows;
            for (int rep = 0; rep < 2; ++rep)
              probe<<<blocks, 64, smem>>>(tm, out, n_box, stages, box_bytes, n_col, n_rowb > 148 ? 148 : n_rowb, box_rows, same);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, out, blocks * 8, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < blocks; ++i) avg += h[i];
            avg /= blocks;
            printf("TMA %-10s box %3d rows (%5d B) stages %d SMs %3d %-9s: %7.1f cyc/box  %6.1f B/cyc/SM  %6.0f ns/box %s\n",
                   strided ? "stride5120" : "contiguous", box_rows, box_bytes, stages, blocks, same ? "same-tile" : "own-tile",
                   avg / n_box, box_bytes / (avg / n_box), avg / n_box / (clk * 1e-6), e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
  return 0;
}
