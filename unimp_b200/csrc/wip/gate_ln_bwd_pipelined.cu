// WORK IN PROGRESS — not built into libunimp_b200.so, never run on a GPU (see README.md here).
//
// K5 backward with the next row prefetched by the bulk-copy engine.
//
// Same arithmetic, same partial-row protocol and the same C-ABI signature as
// unimp_gate_residual_ln_bwd (gate_ln.cu); the difference is how a row reaches the SM:
//   production : every thread issues its 8 LDG.128 at the top of the iteration, then waits
//                (one DRAM round trip per row, nothing overlaps the reduction / store phase);
//   here       : thread 0 of a row group issues up to four `cp.async.bulk` (one per operand row,
//                D*sizeof(T) bytes each) for row i+1 into stage (i+1)&1 BEFORE the group starts
//                on row i; an mbarrier with `complete_tx` signals arrival.  Operands are then read
//                from shared memory (LDS.128), which also frees the 32 staging registers.
// Shared memory per CTA: 2*D floats (column sums) + R * 2 stages * 4 operands * D*sizeof(T)
// (bf16, D = 2560, R = 4: 20 KB + 160 KB).
#include "../common.cuh"
#include "../tc_common.cuh"

namespace unimp {
namespace wip {

using namespace tc;

constexpr int PIPE_MAX_R = 4;
constexpr int PIPE_THREADS = 640;

__device__ __forceinline__ void bulk_load_row(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void group_barrier(int rg, int tg_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
}

__device__ __forceinline__ void group_sum2(float& a, float& b, float* sh, int rg, int tg_threads, int t) {
  const int lane = t & 31, w = t >> 5, nw = tg_threads >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  group_barrier(rg, tg_threads);
  if (lane == 0) { sh[w] = a; sh[32 + w] = b; }
  group_barrier(rg, tg_threads);
  float ra = (lane < nw) ? sh[lane] : 0.f, rb = (lane < nw) ? sh[32 + lane] : 0.f;
  a = warp_sum(ra);
  b = warp_sum(rb);
}

template <typename T, bool COLS, int VPT>
__global__ void __launch_bounds__(PIPE_THREADS)
gate_residual_ln_bwd_pipe_kernel(const T* __restrict__ g_xout, const T* __restrict__ g_ln,
                                 const T* __restrict__ branch, const T* __restrict__ x_out,
                                 const T* __restrict__ gate, const T* __restrict__ gamma,
                                 const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
                                 T* __restrict__ d_x, T* __restrict__ d_branch, float* __restrict__ partial,
                                 int64_t rows, int D, int TG, int R) {
  constexpr int N = Vec16<T>::N;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int nvec = D / N;
  const uint32_t row_bytes = (uint32_t)D * sizeof(T);
  // layout: [2*D floats column sums][R*64 floats scratch][R floats gate][R*2 mbarriers][stages]
  float* colsum = reinterpret_cast<float*>(smem_raw);
  float* scratch = colsum + 2 * D;
  float* sgate = scratch + PIPE_MAX_R * 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sgate + PIPE_MAX_R);
  uint8_t* stages = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(bars + 2 * PIPE_MAX_R) + 127) & ~uintptr_t(127));

  const int rg = threadIdx.x / TG, t = threadIdx.x - rg * TG;
  float* sh = scratch + rg * 64;
  uint64_t* full = bars + 2 * rg;                       // full[stage]
  uint8_t* my_stage = stages + (size_t)rg * 2 * 4 * row_bytes;
  const bool has_ln = g_ln != nullptr && gamma != nullptr;
  const bool gated = branch != nullptr && gate != nullptr;
  const float tg = branch ? (gate ? tanhf(Elem<T>::to_f(*gate)) : 1.f) : 0.f;
  const uint32_t tx_bytes = row_bytes * ((has_ln ? 2u : 0u) + (g_xout ? 1u : 0u) + (gated ? 1u : 0u));

  if (t == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  float dg[COLS ? VPT : 1][N], db[COLS ? VPT : 1][N];
  float dgate = 0.f;
  if (COLS) {
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
      for (int i = 0; i < N; ++i) dg[k][i] = db[k][i] = 0.f;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) colsum[i] = 0.f;
  }
  Vec16<T> gm_raw[VPT];
  if (has_ln) {
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = t + k * TG;
      if (j < nvec) gm_raw[k].load(gamma + j * N);
    }
  }
  __syncthreads();   // barriers initialised, column sums zeroed

  // operand slots of a stage: 0 g_ln, 1 x_out, 2 g_xout, 3 branch
  auto issue = [&](int64_t row, int s) {
    uint8_t* st = my_stage + (size_t)s * 4 * row_bytes;
    mbar_arrive_expect_tx(&full[s], tx_bytes);
    if (has_ln) {
      bulk_load_row(st + 0 * row_bytes, g_ln + row * D, row_bytes, &full[s]);
      bulk_load_row(st + 1 * row_bytes, x_out + row * D, row_bytes, &full[s]);
    }
    if (g_xout) bulk_load_row(st + 2 * row_bytes, g_xout + row * D, row_bytes, &full[s]);
    if (gated) bulk_load_row(st + 3 * row_bytes, branch + row * D, row_bytes, &full[s]);
  };

  const int64_t step = (int64_t)gridDim.x * R;
  int64_t row = (int64_t)blockIdx.x * R + rg;
  if (t == 0 && row < rows) issue(row, 0);
  for (int it = 0; row < rows; row += step, ++it) {
    const int s = it & 1;
    // stage s^1 was last read in iteration it-1, which ended with a group barrier
    if (t == 0 && row + step < rows) issue(row + step, s ^ 1);
    mbar_wait(&full[s], (uint32_t)(it >> 1) & 1u);
    const uint8_t* st = my_stage + (size_t)s * 4 * row_bytes;
    const T* s_gl = reinterpret_cast<const T*>(st);
    const T* s_xo = reinterpret_cast<const T*>(st + row_bytes);
    const T* s_gx = reinterpret_cast<const T*>(st + 2 * row_bytes);
    const T* s_br = reinterpret_cast<const T*>(st + 3 * row_bytes);
    float s1 = 0.f, s2 = 0.f, mean = 0.f, rstd = 0.f;
    if (has_ln) {
      mean = mean_i[row];
      rstd = rstd_i[row];
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int j = t + k * TG;
        if (j < nvec) {
          Vec16<T> a, b;
          float gl[N], xo[N], gm[N];
          a.load(s_gl + j * N);
          b.load(s_xo + j * N);
          gm_raw[k].unpack(gm);
          a.unpack(gl);
          b.unpack(xo);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const float xh = (xo[i] - mean) * rstd, gy = gl[i] * gm[i];
            s1 += gy;
            s2 += gy * xh;
            if (COLS) {
              dg[k][i] += gl[i] * xh;
              db[k][i] += gl[i];
            }
          }
        }
      }
      group_sum2(s1, s2, sh, rg, TG, t);
      s1 /= D;
      s2 /= D;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = t + k * TG;
      if (j < nvec) {
        float dx[N];
        if (g_xout) {
          Vec16<T> a;
          a.load(s_gx + j * N);
          a.unpack(dx);
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) dx[i] = 0.f;
        }
        if (has_ln) {
          Vec16<T> a, b;
          float gl[N], xo[N], gm[N];
          a.load(s_gl + j * N);
          b.load(s_xo + j * N);
          gm_raw[k].unpack(gm);
          a.unpack(gl);
          b.unpack(xo);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const float xh = (xo[i] - mean) * rstd;
            dx[i] += rstd * (gl[i] * gm[i] - s1 - xh * s2);
          }
        }
        Vec16<T> o;
        o.pack(dx);
        o.store(d_x + row * D + j * N);
        if (gated) {
          Vec16<T> bv;
          float bf[N], dbr[N];
          bv.load(s_br + j * N);
          bv.unpack(bf);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            dgate += dx[i] * bf[i];
            dbr[i] = dx[i] * tg;
          }
          Vec16<T> ob;
          ob.pack(dbr);
          ob.store(d_branch + row * D + j * N);
        } else if (d_branch) {
          o.store(d_branch + row * D + j * N);
        }
      }
    }
    group_barrier(rg, TG);   // every thread of the group is done with stage s: it may be refilled
  }

  // fold (identical to the production kernel)
  {
    const int lane = t & 31, w = t >> 5, nw = TG >> 5;
    float v = warp_sum(dgate);
    group_barrier(rg, TG);
    if (lane == 0) sh[w] = v;
    group_barrier(rg, TG);
    float r = (lane < nw) ? sh[lane] : 0.f;
    dgate = warp_sum(r);
  }
  if (t == 0) sgate[rg] = dgate;
  __syncthreads();
  for (int g = 0; COLS && g < R; ++g) {
    if (rg == g && has_ln) {
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int j = t + k * TG;
        if (j < nvec) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            colsum[j * N + i] += dg[k][i];
            colsum[D + j * N + i] += db[k][i];
          }
        }
      }
    }
    __syncthreads();
  }
  float* pr = partial + (int64_t)blockIdx.x * (2 * D + 1);
  if (COLS)
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) pr[i] = colsum[i];
  if (threadIdx.x == 0) {
    float sgsum = 0.f;
    for (int g = 0; g < R; ++g) sgsum += sgate[g];
    pr[2 * D] = sgsum * (1.f - tg * tg);
  }
}

// Geometry: VPT = 2, TG = ceil32(nvec / 2); R row groups such that the stages fit 200 KB.
template <typename T, bool COLS>
static int launch(const void* g_xout, const void* g_ln, const void* branch, const void* x_out, const void* gate,
                  const void* gamma, const float* mean, const float* rstd, void* d_x, void* d_branch,
                  void* partial, int64_t rows, int D, cudaStream_t st, int* G_out) {
  constexpr int N = 16 / (int)sizeof(T);
  const int nvec = D / N;
  const int TG = (((nvec + 1) / 2 + 31) / 32) * 32;
  if (TG > PIPE_THREADS) return -2;
  const size_t row_bytes = (size_t)D * sizeof(T);
  const size_t fixed = (2 * (size_t)D + PIPE_MAX_R * 64 + PIPE_MAX_R) * sizeof(float) + 2 * PIPE_MAX_R * 8 + 128;
  int R = PIPE_THREADS / TG;
  if (R > PIPE_MAX_R) R = PIPE_MAX_R;
  while (R > 0 && fixed + (size_t)R * 2 * 4 * row_bytes > 200 * 1024) --R;
  if (R < 1) return -2;                                  // row too wide for two stages: use production
  const size_t smem = fixed + (size_t)R * 2 * 4 * row_bytes;
  auto kern = gate_residual_ln_bwd_pipe_kernel<T, COLS, 2>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t need = (rows + R - 1) / R;
  const int64_t gmax = UNIMP_NUM_SMS;                    // one CTA per SM (shared memory bound)
  const int64_t passes = (need + gmax - 1) / gmax;
  const int G = (int)((need + passes - 1) / passes);
  kern<<<G, TG * R, smem, st>>>((const T*)g_xout, (const T*)g_ln, (const T*)branch, (const T*)x_out,
                                (const T*)gate, (const T*)gamma, mean, rstd, (T*)d_x, (T*)d_branch,
                                (float*)partial, rows, D, TG, R);
  *G_out = G;
  return 0;
}

}  // namespace wip
}  // namespace unimp

// Main pass only: the caller runs the production fold kernel (gate_residual_ln_bwd_reduce_kernel)
// over partial[G][2D+1] afterwards, exactly as unimp_gate_residual_ln_bwd does.  Returns G (>0) or
// a negative code when the shape does not fit the staged layout.
extern "C" int unimp__gate_residual_ln_bwd_pipelined_main(const void* g_xout, const void* g_ln, const void* branch,
                                                          const void* x_out, const void* gate, const void* gamma,
                                                          const float* mean, const float* rstd, void* d_x,
                                                          void* d_branch, int want_cols, void* partial,
                                                          int64_t rows, int D, int dtype, void* stream) {
  using namespace unimp;
  using namespace unimp::wip;
  if (!d_x || !partial || (!g_xout && !g_ln) || rows <= 0) return -1;
  if ((branch && gate) && !d_branch) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  int G = 0, rc;
  if (dtype == UNIMP_BF16)
    rc = want_cols ? launch<__nv_bfloat16, true>(g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x, d_branch, partial, rows, D, st, &G)
                   : launch<__nv_bfloat16, false>(g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x, d_branch, partial, rows, D, st, &G);
  else
    rc = want_cols ? launch<float, true>(g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x, d_branch, partial, rows, D, st, &G)
                   : launch<float, false>(g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x, d_branch, partial, rows, D, st, &G);
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? G : -(int)e - 100;
}
