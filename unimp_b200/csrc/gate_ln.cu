// K5 — fused tanh-gate + residual + LayerNorm epilogue of GatedCrossAttentionBlock
// (`x = f(x) * tanh(gate) + x`, then the LayerNorm that consumes x).
//
// HBM-bound.  One CTA per row keeps the whole row in registers (<= 4 x 16 B per thread), so
// forward touches every tensor once: read branch, read x, write x_out, write ln_out
// (algorithmic bytes 4*rows*D*sizeof(T)).  Backward is one pass too (read g_xout, g_ln,
// branch, x_out; write d_x, d_branch) with per-CTA column partials for d_gamma/d_beta/d_gate
// reduced by a second small kernel in fixed order (deterministic).
#include "common.cuh"

namespace unimp {

constexpr int LN_VPT = 4;  // 16-byte vectors held per thread

static inline int ln_threads(int D, int n_per_vec) {
  const int nvec = D / n_per_vec;
  int t = (nvec + LN_VPT - 1) / LN_VPT;
  t = ((t + 31) / 32) * 32;
  return t;
}

// SMALL: <= 128 threads per row (D <= 4096 bf16) compiled for <= 56 registers: 96-thread CTAs
// at D = 2560 then sit 12 per SM and all 1536 rows of the 6 x 256 window are resident at once.
template <typename T, bool SMALL>
__global__ void __launch_bounds__(SMALL ? 128 : 1024, SMALL ? 9 : 1) gate_residual_ln_fwd_kernel(
    const T* __restrict__ branch, const T* __restrict__ x, const T* __restrict__ gate,
    const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ x_out,
    T* __restrict__ ln_out, float* __restrict__ mean_o, float* __restrict__ rstd_o, int D, float eps) {
  constexpr int N = Vec16<T>::N;
  const int64_t row = blockIdx.x;
  const int nvec = D / N;
  pdl_launch_dependents();                  // decode steps launch this kernel with a programmatic
                                            // dependency (common.cuh); no-ops otherwise
  // gamma / beta are weights (no kernel of a step writes them).  A decode step (a handful of rows: the
  // launch is one latency chain) pulls this thread's vectors towards L1 now — no registers held — so that
  // the loads after the two reductions do not pay an L2 round trip; with a full grid the extra requests
  // cost more than they save (1536 rows: 6.9 -> 10.9 us, measured).
  if (gamma && gridDim.x <= 64) {
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k) {
      const int j = threadIdx.x + k * blockDim.x;
      if (j < nvec) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(gamma + j * N));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(beta + j * N));
      }
    }
  }
  pdl_wait();
  const T* xr = x + row * D;
  const T* br = branch ? branch + row * D : nullptr;
  const float tg = br ? (gate ? tanhf(Elem<T>::to_f(*gate)) : 1.f) : 0.f;
  // The row stays in registers in its STORAGE type (packed bf16: 4 registers per 8 elements) and
  // is unpacked on demand by each pass; every load of the row is issued before the first use, so
  // a CTA pays one DRAM round trip, and <= 64 registers keep all rows of a 1536 x 2560
  // activation resident in a single wave.
  Vec16<T> xv[LN_VPT], bv[LN_VPT];
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      xv[k].load(xr + j * N);
      if (br) bv[k].load_stream(br + j * N);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      float v[N];
      xv[k].unpack(v);
      if (br) {
        float bf[N];
        bv[k].unpack(bf);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = fmaf(bf[i], tg, v[i]);
        // the residual stream is stored in T: LN must see the rounded value the next
        // consumer reads, so statistics are taken on the rounded x_out.
        xv[k].pack(v);
        xv[k].store(x_out + row * D + j * N);
        xv[k].unpack(v);
      }
#pragma unroll
      for (int i = 0; i < N; ++i) s += v[i];
    }
  }
  if (!gamma) return;
  __shared__ float sh[32];
  const float mean = block_sum(s, sh) / D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      float v[N];
      xv[k].unpack(v);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float d = v[i] - mean;
        q += d * d;
      }
    }
  }
  const float var = block_sum(q, sh) / D;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean_o[row] = mean;
    rstd_o[row] = rstd;
  }
  // gamma / beta are shared by every row (L1/L2 hits); their loads are issued together, after
  // the reductions, so they do not hold registers while the row is in flight
  Vec16<T> gv[LN_VPT], ev[LN_VPT];
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      gv[k].load(gamma + j * N);
      ev[k].load(beta + j * N);
    }
  }
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      float v[N], gf[N], bf[N], o[N];
      xv[k].unpack(v);
      gv[k].unpack(gf);
      ev[k].unpack(bf);
#pragma unroll
      for (int i = 0; i < N; ++i) o[i] = fmaf((v[i] - mean) * rstd, gf[i], bf[i]);
      Vec16<T> ov;
      ov.pack(o);
      ov.store(ln_out + row * D + j * N);
    }
  }
}

// Backward.  R row-groups of TG threads per CTA, each group walking its own rows of a persistent
// grid; a thread owns VPT 16-byte vectors of the row (VPT = 2 for D <= 8192 bf16: TG = 160 at
// D = 2560, every lane busy).  Per row ALL operand loads (g_ln, x_out, g_xout, branch) are issued
// before the first use, so a row costs one DRAM round trip (the previous version loaded g_xout /
// branch only after the row reduction: two dependent round trips per row); gamma sits in
// registers for the whole kernel.  Column partials (d_gamma, d_beta) live in registers per thread,
// are folded across the R groups through shared memory in a fixed order (deterministic), and
// leave the CTA as one partial row: partial[G][2*D + 1] (last = d_gate).
// Ungated residual (gate == NULL): d_branch == d_x, so `branch` is not read and d_branch is only
// written if the caller asks for it (ops.py hands d_x to both inputs): 4 passes instead of 6.
constexpr int LN_BWD_MAX_R = 12;       // named barriers 1..12
constexpr int LN_BWD_MAX_THREADS = 640;

__device__ __forceinline__ float group_sum(float v, float* sh, int rg, int tg_threads, int t) {
  const int lane = t & 31, w = t >> 5, nw = tg_threads >> 5;
  v = warp_sum(v);
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  if (lane == 0) sh[w] = v;
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  float r = (lane < nw) ? sh[lane] : 0.f;
  return warp_sum(r);
}

// two sums with one pair of barriers (sh: 64 floats per group, up to 32 warps per group)
__device__ __forceinline__ void group_sum2(float& a, float& b, float* sh, int rg, int tg_threads, int t) {
  const int lane = t & 31, w = t >> 5, nw = tg_threads >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  if (lane == 0) { sh[w] = a; sh[32 + w] = b; }
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  float ra = (lane < nw) ? sh[lane] : 0.f, rb = (lane < nw) ? sh[32 + lane] : 0.f;
  a = warp_sum(ra);
  b = warp_sum(rb);
}

// COLS = false: the LayerNorm's affine parameters are frozen (the LM / ViT towers): no column
// partials are kept.
template <typename T, bool COLS, int VPT>
__global__ void __launch_bounds__(LN_BWD_MAX_THREADS)
gate_residual_ln_bwd_kernel(const T* __restrict__ g_xout, const T* __restrict__ g_ln,
                            const T* __restrict__ branch, const T* __restrict__ x_out,
                            const T* __restrict__ gate, const T* __restrict__ gamma,
                            const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
                            T* __restrict__ d_x, T* __restrict__ d_branch, float* __restrict__ partial,
                            int64_t rows, int D, int TG, int R) {
  constexpr int N = Vec16<T>::N;
  extern __shared__ float sdyn[];          // [2*D] column sums, then MAX_R*64 scratch, then MAX_R
  const int nvec = D / N;
  const int rg = threadIdx.x / TG, t = threadIdx.x - rg * TG;
  float* sh = sdyn + 2 * D + rg * 64;
  float* sgate = sdyn + 2 * D + LN_BWD_MAX_R * 64;
  const bool has_ln = g_ln != nullptr && gamma != nullptr;
  const bool gated = branch != nullptr && gate != nullptr;   // only then is `branch` needed
  const float tg = branch ? (gate ? tanhf(Elem<T>::to_f(*gate)) : 1.f) : 0.f;
  float dg[COLS ? VPT : 1][N], db[COLS ? VPT : 1][N];
  float dgate = 0.f;
  if (COLS) {
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
      for (int i = 0; i < N; ++i) dg[k][i] = db[k][i] = 0.f;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sdyn[i] = 0.f;
  }
  Vec16<T> gm_raw[VPT];
  if (has_ln) {
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = t + k * TG;
      if (j < nvec) gm_raw[k].load(gamma + j * N);
    }
  }

  for (int64_t row = (int64_t)blockIdx.x * R + rg; row < rows; row += (int64_t)gridDim.x * R) {
    Vec16<T> gl_raw[VPT], xo_raw[VPT], gx_raw[VPT], br_raw[VPT];
    float s1 = 0.f, s2 = 0.f, mean = 0.f, rstd = 0.f;
    const int64_t base = row * D;
    // ---- every load of the row, up front -------------------------------------------------
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int j = t + k * TG;
      if (j < nvec) {
        if (has_ln) {
          gl_raw[k].load_stream(g_ln + base + j * N);
          xo_raw[k].load_stream(x_out + base + j * N);
        }
        if (g_xout) gx_raw[k].load_stream(g_xout + base + j * N);
        if (gated) br_raw[k].load_stream(branch + base + j * N);
      }
    }
    if (has_ln) {
      mean = mean_i[row];
      rstd = rstd_i[row];
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int j = t + k * TG;
        if (j < nvec) {
          float gl[N], xo[N], gm[N];
          gm_raw[k].unpack(gm);
          gl_raw[k].unpack(gl);
          xo_raw[k].unpack(xo);
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
          gx_raw[k].unpack(dx);
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) dx[i] = 0.f;
        }
        if (has_ln) {
          float gl[N], xo[N], gm[N];
          gm_raw[k].unpack(gm);
          gl_raw[k].unpack(gl);
          xo_raw[k].unpack(xo);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const float xh = (xo[i] - mean) * rstd;
            dx[i] += rstd * (gl[i] * gm[i] - s1 - xh * s2);
          }
        }
        Vec16<T> o;
        o.pack(dx);
        o.store(d_x + base + j * N);
        if (gated) {
          float bf[N], dbr[N];
          br_raw[k].unpack(bf);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            dgate += dx[i] * bf[i];
            dbr[i] = dx[i] * tg;
          }
          Vec16<T> ob;
          ob.pack(dbr);
          ob.store(d_branch + base + j * N);
        } else if (d_branch) {
          o.store(d_branch + base + j * N);   // ungated residual: d_branch == d_x
        }
      }
    }
  }
  // fold the R groups' column partials in a fixed order (deterministic), then emit one row
  dgate = group_sum(dgate, sh, rg, TG, t);
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
            sdyn[j * N + i] += dg[k][i];
            sdyn[D + j * N + i] += db[k][i];
          }
        }
      }
    }
    __syncthreads();
  }
  float* pr = partial + (int64_t)blockIdx.x * (2 * D + 1);
  if (COLS)
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) pr[i] = sdyn[i];
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int g = 0; g < R; ++g) s += sgate[g];
    pr[2 * D] = s * (1.f - tg * tg);
  }
}

// Reduce partial[G][2D+1] over G in a fixed order: block = 32 columns x 8 g-lanes.
// accumulate bit 0: d_gate += ; bit 1: d_gamma / d_beta += (the outputs are the optimizer's gradient
// buffers and an earlier micro-batch of the step has already written them)
template <typename T>
__global__ void gate_residual_ln_bwd_reduce_kernel(const float* __restrict__ partial, int G, int D,
                                                   T* __restrict__ d_gate, T* __restrict__ d_gamma,
                                                   T* __restrict__ d_beta, int accumulate) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int W = 2 * D + 1;
  float s = 0.f;
  if (c < W)
    for (int g = gy; g < G; g += 8) s += partial[(int64_t)g * W + c];
  sh[gy][cx] = s;
  __syncthreads();
  if (gy == 0 && c < W) {
    s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += sh[g][cx];
    if (c < D) {
      if (d_gamma) d_gamma[c] = Elem<T>::from_f(s + ((accumulate & 2) ? Elem<T>::to_f(d_gamma[c]) : 0.f));
    } else if (c < 2 * D) {
      if (d_beta) d_beta[c - D] = Elem<T>::from_f(s + ((accumulate & 2) ? Elem<T>::to_f(d_beta[c - D]) : 0.f));
    } else {
      if (d_gate) d_gate[0] = Elem<T>::from_f(s + ((accumulate & 1) ? Elem<T>::to_f(d_gate[0]) : 0.f));
    }
  }
}

// Launch geometry of the backward: VPT, threads per row group, row groups per CTA, grid.
struct LnBwdGeom {
  int vpt, TG, R, G, smem;
};

template <typename K>
static int ln_bwd_ctas_per_sm(K kernel, int threads, int smem) {
  int n = 0;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  return n;
}

template <typename T, bool COLS>
static LnBwdGeom ln_bwd_geom(int64_t rows, int D) {
  constexpr int N = 16 / (int)sizeof(T);
  const int nvec = D / N;
  LnBwdGeom g;
  int tg2 = (((nvec + 1) / 2 + 31) / 32) * 32;
  if (tg2 <= 512) {
    g.vpt = 2;
    g.TG = tg2;
  } else {                       // wide rows (D > 8192 bf16): correctness path, spills
    g.vpt = 4;
    g.TG = (((nvec + 3) / 4 + 31) / 32) * 32;   // <= 512: the entry point bounds D
  }
  // column partials cost one partial row per CTA: few, fat CTAs (one per SM).  Without them,
  // 320-thread CTAs (two or three per SM) keep more rows in flight.
  const int target = COLS ? LN_BWD_MAX_THREADS : 320;
  int R = target / g.TG;
  R = R < 1 ? 1 : (R > LN_BWD_MAX_R ? LN_BWD_MAX_R : R);
  g.R = R;
  const int threads = g.TG * R;
  g.smem = (2 * D + LN_BWD_MAX_R * 64 + LN_BWD_MAX_R) * (int)sizeof(float);
  // resident CTAs per SM: asked of the runtime once per (kernel, threads, smem)
  constexpr int NC = 16;
  static int cached_key[NC], cached_val[NC], n_cached = 0;
  const int key = (g.vpt << 28) ^ (threads << 17) ^ g.smem;
  int per_sm = 0;
  for (int i = 0; i < n_cached; ++i)
    if (cached_key[i] == key) per_sm = cached_val[i];
  if (!per_sm) {
    per_sm = g.vpt == 2
        ? ln_bwd_ctas_per_sm(gate_residual_ln_bwd_kernel<T, COLS, 2>, threads, g.smem)
        : ln_bwd_ctas_per_sm(gate_residual_ln_bwd_kernel<T, COLS, 4>, threads, g.smem);
    if (n_cached < NC) {
      cached_key[n_cached] = key;
      cached_val[n_cached] = per_sm;
      ++n_cached;
    }
  }
  int64_t gmax = (int64_t)per_sm * UNIMP_NUM_SMS;
  if (gmax > 3 * UNIMP_NUM_SMS) gmax = 3 * UNIMP_NUM_SMS;   // workspace holds 3*SMs partial rows
  // balanced passes: every CTA walks the same number of rows (no ragged last wave)
  const int64_t need = (rows + R - 1) / R;
  const int64_t passes = (need + gmax - 1) / gmax;
  g.G = (int)((need + passes - 1) / passes);
  return g;
}

template <typename T, bool COLS>
static void ln_bwd_launch(const LnBwdGeom& g, cudaStream_t st, const void* g_xout, const void* g_ln,
                          const void* branch, const void* x_out, const void* gate, const void* gamma,
                          const float* mean, const float* rstd, void* d_x, void* d_branch,
                          void* partial, int64_t rows, int D) {
  if (g.vpt == 2)
    gate_residual_ln_bwd_kernel<T, COLS, 2><<<g.G, g.TG * g.R, g.smem, st>>>(
        (const T*)g_xout, (const T*)g_ln, (const T*)branch, (const T*)x_out, (const T*)gate,
        (const T*)gamma, mean, rstd, (T*)d_x, (T*)d_branch, (float*)partial, rows, D, g.TG, g.R);
  else
    gate_residual_ln_bwd_kernel<T, COLS, 4><<<g.G, g.TG * g.R, g.smem, st>>>(
        (const T*)g_xout, (const T*)g_ln, (const T*)branch, (const T*)x_out, (const T*)gate,
        (const T*)gamma, mean, rstd, (T*)d_x, (T*)d_branch, (float*)partial, rows, D, g.TG, g.R);
}

}  // namespace unimp

using namespace unimp;

extern "C" int unimp_gate_residual_ln_fwd(const void* branch, const void* x, const void* gate,
                                          const void* gamma, const void* beta, void* x_out,
                                          void* ln_out, float* mean, float* rstd, int64_t rows,
                                          int D, float eps, int dtype, void* stream) {
  UNIMP_CHECK_ARG(x, UNIMP_E_NULL, "gate_residual_ln_fwd: x is NULL");
  UNIMP_CHECK_ARG(!branch || x_out, UNIMP_E_NULL, "gate_residual_ln_fwd: branch given without x_out");
  UNIMP_CHECK_ARG(!gamma || (beta && ln_out && mean && rstd), UNIMP_E_NULL,
                  "gate_residual_ln_fwd: gamma given without beta/ln_out/mean/rstd");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE,
                  "gate_residual_ln_fwd: dtype");
  const int npv = dtype == UNIMP_BF16 ? 8 : 4;
  UNIMP_CHECK_ARG(rows >= 0 && D > 0 && D % npv == 0 && D / npv <= 1024 * LN_VPT, UNIMP_E_SHAPE,
                  "gate_residual_ln_fwd: D=%d must be a multiple of %d and <= %d", D, npv,
                  1024 * LN_VPT * npv);
  UNIMP_CHECK_ARG(aligned16(x) && aligned16(branch) && aligned16(gamma) && aligned16(beta) &&
                      aligned16(x_out) && aligned16(ln_out),
                  UNIMP_E_ALIGN, "gate_residual_ln_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const int threads = ln_threads(D, npv);
  cudaStream_t st = (cudaStream_t)stream;
  // a decode step (a handful of rows, ~150 of these launches per token) is launched with a programmatic
  // dependency so that it is scheduled while its predecessor drains
#define UNIMP_LN_FWD_LAUNCH(TT, SM)                                                                       \
  do {                                                                                                    \
    if (rows <= 64) {                                                                                     \
      cudaError_t e_ = launch_pdl(gate_residual_ln_fwd_kernel<TT, SM>, dim3((unsigned)rows), dim3(threads), 0, \
                                  st, (const TT*)branch, (const TT*)x, (const TT*)gate, (const TT*)gamma,  \
                                  (const TT*)beta, (TT*)x_out, (TT*)ln_out, mean, rstd, D, eps);           \
      if (e_ != cudaSuccess) { set_error("gate_residual_ln_fwd launch: %s", cudaGetErrorString(e_)); return (int)e_; } \
    } else {                                                                                              \
      gate_residual_ln_fwd_kernel<TT, SM><<<(unsigned)rows, threads, 0, st>>>(                             \
          (const TT*)branch, (const TT*)x, (const TT*)gate, (const TT*)gamma, (const TT*)beta, (TT*)x_out, \
          (TT*)ln_out, mean, rstd, D, eps);                                                               \
    }                                                                                                     \
  } while (0)
  if (dtype == UNIMP_BF16) {
    if (threads <= 128) UNIMP_LN_FWD_LAUNCH(__nv_bfloat16, true);
    else UNIMP_LN_FWD_LAUNCH(__nv_bfloat16, false);
  } else {
    if (threads <= 128) UNIMP_LN_FWD_LAUNCH(float, true);
    else UNIMP_LN_FWD_LAUNCH(float, false);
  }
#undef UNIMP_LN_FWD_LAUNCH
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t unimp_gate_residual_ln_bwd_workspace(int64_t rows, int D) {
  (void)rows;
  return 3 * (int64_t)UNIMP_NUM_SMS * (2 * (int64_t)D + 1) * sizeof(float);
}

extern "C" int unimp_gate_residual_ln_bwd(const void* g_xout, const void* g_ln, const void* branch,
                                          const void* x_out, const void* gate, const void* gamma,
                                          const float* mean, const float* rstd, void* d_x,
                                          void* d_branch, void* d_gate, void* d_gamma,
                                          void* d_beta, void* partial, int64_t rows, int D,
                                          int accumulate, int dtype, void* stream) {
  UNIMP_CHECK_ARG(d_x && partial, UNIMP_E_NULL, "gate_residual_ln_bwd: d_x/partial NULL");
  UNIMP_CHECK_ARG(g_xout || g_ln, UNIMP_E_NULL, "gate_residual_ln_bwd: no incoming gradient");
  UNIMP_CHECK_ARG(!g_ln || (gamma && x_out && mean && rstd), UNIMP_E_NULL,
                  "gate_residual_ln_bwd: g_ln given without gamma/x_out/mean/rstd");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE,
                  "gate_residual_ln_bwd: dtype");
  const int npv = dtype == UNIMP_BF16 ? 8 : 4;
  UNIMP_CHECK_ARG(rows > 0 && D > 0 && D % npv == 0 && D / npv <= 2048, UNIMP_E_SHAPE,
                  "gate_residual_ln_bwd: bad rows/D (D=%d must be a multiple of %d and <= %d)", D, npv,
                  2048 * npv);
  UNIMP_CHECK_ARG(aligned16(g_xout) && aligned16(g_ln) && aligned16(branch) && aligned16(x_out) &&
                      aligned16(gamma) && aligned16(d_x) && aligned16(d_branch),
                  UNIMP_E_ALIGN, "gate_residual_ln_bwd: pointers must be 16-byte aligned");
  UNIMP_CHECK_ARG(!(branch && gate) || d_branch, UNIMP_E_NULL,
                  "gate_residual_ln_bwd: gated branch given without d_branch");
  cudaStream_t st = (cudaStream_t)stream;
  const bool cols = (d_gamma || d_beta) && g_ln;
  LnBwdGeom g;
  if (dtype == UNIMP_BF16) {
    if (cols) {
      g = ln_bwd_geom<__nv_bfloat16, true>(rows, D);
      ln_bwd_launch<__nv_bfloat16, true>(g, st, g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x,
                                         d_branch, partial, rows, D);
    } else {
      g = ln_bwd_geom<__nv_bfloat16, false>(rows, D);
      ln_bwd_launch<__nv_bfloat16, false>(g, st, g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x,
                                          d_branch, partial, rows, D);
    }
  } else {
    if (cols) {
      g = ln_bwd_geom<float, true>(rows, D);
      ln_bwd_launch<float, true>(g, st, g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x,
                                 d_branch, partial, rows, D);
    } else {
      g = ln_bwd_geom<float, false>(rows, D);
      ln_bwd_launch<float, false>(g, st, g_xout, g_ln, branch, x_out, gate, gamma, mean, rstd, d_x,
                                  d_branch, partial, rows, D);
    }
  }
  const int G = g.G;
  UNIMP_CHECK_LAUNCH();
  if (!cols && !d_gate) return 0;
  const int W = 2 * D + 1;
  if (dtype == UNIMP_BF16)
    gate_residual_ln_bwd_reduce_kernel<__nv_bfloat16><<<(W + 31) / 32, 256, 0, st>>>(
        (const float*)partial, G, D, (__nv_bfloat16*)d_gate, (__nv_bfloat16*)d_gamma,
        (__nv_bfloat16*)d_beta, accumulate);
  else
    gate_residual_ln_bwd_reduce_kernel<float><<<(W + 31) / 32, 256, 0, st>>>(
        (const float*)partial, G, D, (float*)d_gate, (float*)d_gamma, (float*)d_beta, accumulate);
  UNIMP_CHECK_LAUNCH();
  return 0;
}
