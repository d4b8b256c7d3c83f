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

template <typename T>
__global__ void gate_residual_ln_fwd_kernel(const T* __restrict__ branch, const T* __restrict__ x,
                                            const T* __restrict__ gate,
                                            const T* __restrict__ gamma, const T* __restrict__ beta,
                                            T* __restrict__ x_out, T* __restrict__ ln_out,
                                            float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                            int D, float eps) {
  constexpr int N = Vec16<T>::N;
  const int64_t row = blockIdx.x;
  const int nvec = D / N;
  const T* xr = x + row * D;
  const T* br = branch ? branch + row * D : nullptr;
  float v[LN_VPT][N];
  const float tg = br ? (gate ? tanhf(Elem<T>::to_f(*gate)) : 1.f) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      Vec16<T> a;
      a.load(xr + j * N);
      a.unpack(v[k]);
      if (br) {
        Vec16<T> b;
        float bf[N];
        b.load_stream(br + j * N);
        b.unpack(bf);
#pragma unroll
        for (int i = 0; i < N; ++i) v[k][i] = fmaf(bf[i], tg, v[k][i]);
        // the residual stream is stored in T: LN must see the rounded value the next
        // consumer reads, so statistics are taken on the rounded x_out.
        Vec16<T> o;
        o.pack(v[k]);
        o.store(x_out + row * D + j * N);
        o.unpack(v[k]);
      }
#pragma unroll
      for (int i = 0; i < N; ++i) s += v[k][i];
    }
  }
  if (!gamma) return;
  // issue the affine-parameter loads now so their latency hides under the two reductions
  Vec16<T> gv[LN_VPT], bv[LN_VPT];
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      gv[k].load(gamma + j * N);
      bv[k].load(beta + j * N);
    }
  }
  __shared__ float sh[32];
  const float mean = block_sum(s, sh) / D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float d = v[k][i] - mean;
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
#pragma unroll
  for (int k = 0; k < LN_VPT; ++k) {
    const int j = threadIdx.x + k * blockDim.x;
    if (j < nvec) {
      float gf[N], bf[N], o[N];
      gv[k].unpack(gf);
      bv[k].unpack(bf);
#pragma unroll
      for (int i = 0; i < N; ++i) o[i] = fmaf((v[k][i] - mean) * rstd, gf[i], bf[i]);
      Vec16<T> ov;
      ov.pack(o);
      ov.store(ln_out + row * D + j * N);
    }
  }
}

// R row-groups of TG threads per CTA, each group walking its own rows, so that (nearly) every row
// of a 768-row activation is in flight at once; column partials (d_gamma, d_beta) live in
// registers per thread, are folded across the R groups through shared memory in a fixed order,
// and leave the CTA as one partial row: partial[G][2*D + 1] (last = d_gate).
constexpr int LN_BWD_MAX_R = 12;       // row-groups per CTA: R = clamp(192 / TG, 1, 12): small CTAs,
                                       // so 2-3 of them fit the register file of an SM
static inline int ln_bwd_r(int TG) { int r = 192 / TG; return r < 1 ? 1 : (r > LN_BWD_MAX_R ? LN_BWD_MAX_R : r); }

__device__ __forceinline__ float group_sum(float v, float* sh, int rg, int tg_threads, int t) {
  const int lane = t & 31, w = t >> 5, nw = tg_threads >> 5;
  v = warp_sum(v);
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  if (lane == 0) sh[w] = v;
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  float r = (lane < nw) ? sh[lane] : 0.f;
  return warp_sum(r);
}

// COLS = false: the LayerNorm's affine parameters are frozen (the LM / ViT towers): no column
// partials are kept, which frees 64 registers per thread and the final fold.
// two sums with one pair of barriers
__device__ __forceinline__ void group_sum2(float& a, float& b, float* sh, int rg, int tg_threads, int t) {
  const int lane = t & 31, w = t >> 5, nw = tg_threads >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  if (lane == 0) { sh[w] = a; sh[16 + w] = b; }
  asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(tg_threads) : "memory");
  float ra = (lane < nw) ? sh[lane] : 0.f, rb = (lane < nw) ? sh[16 + lane] : 0.f;
  a = warp_sum(ra);
  b = warp_sum(rb);
}

template <typename T, bool COLS>
__global__ void gate_residual_ln_bwd_kernel(const T* __restrict__ g_xout, const T* __restrict__ g_ln,
                                            const T* __restrict__ branch, const T* __restrict__ x_out,
                                            const T* __restrict__ gate,
                                            const T* __restrict__ gamma,
                                            const float* __restrict__ mean_i,
                                            const float* __restrict__ rstd_i, T* __restrict__ d_x,
                                            T* __restrict__ d_branch, float* __restrict__ partial,
                                            int64_t rows, int D, int TG, int R) {
  constexpr int N = Vec16<T>::N;
  extern __shared__ float sdyn[];          // [2*D] column sums, then LN_BWD_R*32 scratch, then R
  const int nvec = D / N;
  const int rg = threadIdx.x / TG, t = threadIdx.x - rg * TG;
  float* sh = sdyn + 2 * D + rg * 32;
  float* sgate = sdyn + 2 * D + LN_BWD_MAX_R * 32;
  const bool has_ln = g_ln != nullptr && gamma != nullptr;
  const float tg = branch ? (gate ? tanhf(Elem<T>::to_f(*gate)) : 1.f) : 0.f;
  float dg[COLS ? LN_VPT : 1][N], db[COLS ? LN_VPT : 1][N];
  float dgate = 0.f;
  if (COLS) {
#pragma unroll
    for (int k = 0; k < LN_VPT; ++k)
#pragma unroll
      for (int i = 0; i < N; ++i) dg[k][i] = db[k][i] = 0.f;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sdyn[i] = 0.f;
  }

  for (int64_t row = (int64_t)blockIdx.x * R + rg; row < rows; row += (int64_t)gridDim.x * R) {
    Vec16<T> gl_raw[LN_VPT], xo_raw[LN_VPT];
    float s1 = 0.f, s2 = 0.f, mean = 0.f, rstd = 0.f;
    if (has_ln) {
      mean = mean_i[row];
      rstd = rstd_i[row];
#pragma unroll
      for (int k = 0; k < LN_VPT; ++k) {
        const int j = t + k * TG;
        if (j < nvec) {
          gl_raw[k].load_stream(g_ln + row * D + j * N);
          xo_raw[k].load_stream(x_out + row * D + j * N);
        }
      }
#pragma unroll
      for (int k = 0; k < LN_VPT; ++k) {
        const int j = t + k * TG;
        if (j < nvec) {
          Vec16<T> gmv;
          float gl[N], xo[N], gm[N];
          gmv.load(gamma + j * N);
          gmv.unpack(gm);
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
    for (int k = 0; k < LN_VPT; ++k) {
      const int j = t + k * TG;
      if (j < nvec) {
        float dx[N];
        if (g_xout) {
          Vec16<T> a;
          a.load_stream(g_xout + row * D + j * N);
          a.unpack(dx);
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) dx[i] = 0.f;
        }
        if (has_ln) {
          Vec16<T> gmv;
          float gl[N], xo[N], gm[N];
          gmv.load(gamma + j * N);
          gmv.unpack(gm);
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
        o.store(d_x + row * D + j * N);
        if (branch) {
          Vec16<T> b;
          float bf[N], dbr[N];
          b.load_stream(branch + row * D + j * N);
          b.unpack(bf);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            dgate += dx[i] * bf[i];
            dbr[i] = dx[i] * tg;
          }
          Vec16<T> ob;
          ob.pack(dbr);
          ob.store(d_branch + row * D + j * N);
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
      for (int k = 0; k < LN_VPT; ++k) {
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
template <typename T>
__global__ void gate_residual_ln_bwd_reduce_kernel(const float* __restrict__ partial, int G, int D,
                                                   T* __restrict__ d_gate, T* __restrict__ d_gamma,
                                                   T* __restrict__ d_beta) {
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
      if (d_gamma) d_gamma[c] = Elem<T>::from_f(s);
    } else if (c < 2 * D) {
      if (d_beta) d_beta[c - D] = Elem<T>::from_f(s);
    } else {
      if (d_gate) d_gate[0] = Elem<T>::from_f(s);
    }
  }
}

static inline int ln_bwd_grid(int64_t rows, int R, bool cols) {
  const int64_t need = (rows + R - 1) / R;
  // 192-thread CTAs: ~150 registers with column partials (2 CTAs/SM), ~96 without (3 CTAs/SM)
  const int64_t g = cols ? 2 * UNIMP_NUM_SMS : 3 * UNIMP_NUM_SMS;
  return (int)(need < g ? need : g);
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
  if (dtype == UNIMP_BF16)
    gate_residual_ln_fwd_kernel<__nv_bfloat16><<<(unsigned)rows, threads, 0, st>>>(
        (const __nv_bfloat16*)branch, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gate,
        (const __nv_bfloat16*)gamma,
        (const __nv_bfloat16*)beta, (__nv_bfloat16*)x_out, (__nv_bfloat16*)ln_out, mean, rstd, D,
        eps);
  else
    gate_residual_ln_fwd_kernel<float><<<(unsigned)rows, threads, 0, st>>>(
        (const float*)branch, (const float*)x, (const float*)gate, (const float*)gamma,
        (const float*)beta,
        (float*)x_out, (float*)ln_out, mean, rstd, D, eps);
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
                                          int dtype, void* stream) {
  UNIMP_CHECK_ARG(d_x && partial, UNIMP_E_NULL, "gate_residual_ln_bwd: d_x/partial NULL");
  UNIMP_CHECK_ARG(g_xout || g_ln, UNIMP_E_NULL, "gate_residual_ln_bwd: no incoming gradient");
  UNIMP_CHECK_ARG(!g_ln || (gamma && x_out && mean && rstd), UNIMP_E_NULL,
                  "gate_residual_ln_bwd: g_ln given without gamma/x_out/mean/rstd");
  UNIMP_CHECK_ARG(!branch || d_branch, UNIMP_E_NULL,
                  "gate_residual_ln_bwd: branch given without d_branch");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE,
                  "gate_residual_ln_bwd: dtype");
  const int npv = dtype == UNIMP_BF16 ? 8 : 4;
  UNIMP_CHECK_ARG(rows > 0 && D > 0 && D % npv == 0 && D / npv <= 1024 * LN_VPT, UNIMP_E_SHAPE,
                  "gate_residual_ln_bwd: bad rows/D");
  UNIMP_CHECK_ARG(aligned16(g_xout) && aligned16(g_ln) && aligned16(branch) && aligned16(x_out) &&
                      aligned16(gamma) && aligned16(d_x) && aligned16(d_branch),
                  UNIMP_E_ALIGN, "gate_residual_ln_bwd: pointers must be 16-byte aligned");
  const int TG = ln_threads(D, npv);
  const int R = ln_bwd_r(TG);
  UNIMP_CHECK_ARG(TG * R <= 384, UNIMP_E_SHAPE,
                  "gate_residual_ln_bwd: D=%d too large for the register budget", D);
  const int threads = TG * R;
  const int smem = (2 * D + LN_BWD_MAX_R * 32 + LN_BWD_MAX_R) * (int)sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const bool cols = (d_gamma || d_beta) && g_ln;
  const int G = ln_bwd_grid(rows, R, cols);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gate_residual_ln_bwd_kernel<__nv_bfloat16, true>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gate_residual_ln_bwd_kernel<float, true>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gate_residual_ln_bwd_kernel<__nv_bfloat16, false>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gate_residual_ln_bwd_kernel<float, false>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
#define UNIMP_LN_BWD_LAUNCH(TT, CC)                                                                  \
  gate_residual_ln_bwd_kernel<TT, CC><<<G, threads, smem, st>>>(                                      \
      (const TT*)g_xout, (const TT*)g_ln, (const TT*)branch, (const TT*)x_out, (const TT*)gate,       \
      (const TT*)gamma, mean, rstd, (TT*)d_x, (TT*)d_branch, (float*)partial, rows, D, TG, R)
  if (dtype == UNIMP_BF16) {
    if (cols) UNIMP_LN_BWD_LAUNCH(__nv_bfloat16, true);
    else UNIMP_LN_BWD_LAUNCH(__nv_bfloat16, false);
  } else {
    if (cols) UNIMP_LN_BWD_LAUNCH(float, true);
    else UNIMP_LN_BWD_LAUNCH(float, false);
  }
#undef UNIMP_LN_BWD_LAUNCH
  UNIMP_CHECK_LAUNCH();
  if (!cols && !d_gate) return 0;
  const int W = 2 * D + 1;
  if (dtype == UNIMP_BF16)
    gate_residual_ln_bwd_reduce_kernel<__nv_bfloat16><<<(W + 31) / 32, 256, 0, st>>>(
        (const float*)partial, G, D, (__nv_bfloat16*)d_gate, (__nv_bfloat16*)d_gamma,
        (__nv_bfloat16*)d_beta);
  else
    gate_residual_ln_bwd_reduce_kernel<float><<<(W + 31) / 32, 256, 0, st>>>(
        (const float*)partial, G, D, (float*)d_gate, (float*)d_gamma, (float*)d_beta);
  UNIMP_CHECK_LAUNCH();
  return 0;
}
