// Elementwise fusions around the frozen towers (SURVEY.md §8f-3 "next" row; K4 is cuDNN SDPA):
//   unimp_rotary_qkv_fwd/bwd : GPT-NeoX rotary embedding over the packed (B,T,H,3*dh) output of
//       query_key_value into one packed buffer, so q/k/v reach SDPA as strided views with no
//       chunk / cat / rotate_half temporaries (HF apply_rotary_pos_emb: ~10 launches -> 1);
//       backward un-rotates dq/dk and packs dq|dk|dv into d_qkv in one pass.
//   unimp_quick_gelu : x * sigmoid(1.702 x), in place (CLIP ViT MLP, forward only).
//   unimp_gelu_fwd/bwd : exact (erf) GELU of the FeedForward blocks (GatedCrossAttentionBlock.ff,
//       PerceiverResampler ff, GPT-NeoX mlp.act).  HBM-bound by design: 4 vectors in flight per
//       thread; in bf16 the normal CDF comes from a 1.5e-7-accurate rational erfc (2 MUFU + ~12
//       FP32 ops per element instead of erff's ~25, which is what keeps the stock kernel
//       issue-bound at half the HBM rate); fp32 (the 1e-4 parity mode) uses erff itself.
#include "common.cuh"

namespace unimp {

// out[b,t,h,{q,k,v},:] = rotary(qkv) (v and the non-rotary tail copied); thread = one 16-byte
// vector.  cos/sin: (cb, T, rot) with the two halves duplicated (HF layout).
template <typename T>
__global__ void rotary_qkv_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                      const T* __restrict__ cs, const T* __restrict__ sn, int64_t total,
                                      int Tn, int H, int dh, int rot, int64_t cs_bstride) {
  constexpr int N = Vec16<T>::N;
  const int half = rot / 2, vpd = dh / N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vpd);
  const int which = (int)((idx / vpd) % 3);
  const int64_t bth = idx / (3 * vpd);
  const int64_t bt = bth / H;
  const int t = (int)(bt % Tn);
  const int64_t b = bt / Tn;
  const int d0 = v * N;
  const T* src = qkv + bth * 3 * dh + which * dh;
  T* dst = out + bth * 3 * dh + which * dh + d0;
  Vec16<T> x;
  x.load(src + d0);
  if (which == 2 || d0 >= rot) {
    x.store(dst);
    return;
  }
  // q*cos + rotate_half(q)*sin: first half: x1 c - x2 s ; second half: x2 c + x1 s
  const bool first = d0 < half;
  Vec16<T> xo, c, s_;
  xo.load(src + (first ? d0 + half : d0 - half));
  c.load(cs + b * cs_bstride + (int64_t)t * rot + d0);
  s_.load(sn + b * cs_bstride + (int64_t)t * rot + d0);
  float xf[N], xof[N], cf[N], sf[N], o[N];
  x.unpack(xf); xo.unpack(xof); c.unpack(cf); s_.unpack(sf);
#pragma unroll
  for (int i = 0; i < N; ++i) o[i] = first ? xf[i] * cf[i] - xof[i] * sf[i] : xf[i] * cf[i] + xof[i] * sf[i];
  Vec16<T> ov;
  ov.pack(o);
  ov.store(dst);
}

// d_qkv[b,t,h,{q,k,v},:] from dq/dk/dv (B,H,T,dh) strided views; thread = one 16-byte vector.
template <typename T>
__global__ void rotary_qkv_bwd_kernel(const T* __restrict__ dq, const T* __restrict__ dk,
                                      const T* __restrict__ dv, int64_t q_sb, int64_t q_sh, int64_t q_st,
                                      int64_t k_sb, int64_t k_sh, int64_t k_st, int64_t v_sb, int64_t v_sh,
                                      int64_t v_st, const T* __restrict__ cs, const T* __restrict__ sn,
                                      T* __restrict__ d_qkv, int64_t total, int Tn, int H, int dh, int rot,
                                      int64_t cs_bstride) {
  constexpr int N = Vec16<T>::N;
  const int half = rot / 2, vpd = dh / N;  // vectors per head-dim
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vpd);
  const int which = (int)((idx / vpd) % 3);
  const int64_t bth = idx / (3 * vpd);
  const int h = (int)(bth % H);
  const int64_t bt = bth / H;
  const int t = (int)(bt % Tn);
  const int64_t b = bt / Tn;
  const T* src = which == 0 ? dq + b * q_sb + h * q_sh + (int64_t)t * q_st
               : which == 1 ? dk + b * k_sb + h * k_sh + (int64_t)t * k_st
                            : dv + b * v_sb + h * v_sh + (int64_t)t * v_st;
  T* dst = d_qkv + bth * 3 * dh + which * dh + v * N;
  const int d0 = v * N;
  Vec16<T> g;
  g.load(src + d0);
  if (which == 2 || d0 >= rot) {
    g.store(dst);
    return;
  }
  // out1 = x1 c1 - x2 s1 ; out2 = x2 c2 + x1 s2   =>   dx1 = g1 c1 + g2 s2 ; dx2 = -g1 s1 + g2 c2
  const bool first = d0 < half;
  Vec16<T> go, c, s;
  go.load(src + (first ? d0 + half : d0 - half));
  const T* cp = cs + b * cs_bstride + (int64_t)t * rot;
  const T* sp = sn + b * cs_bstride + (int64_t)t * rot;
  float gf[N], gof[N], cf[N], sf[N], o[N];
  g.unpack(gf); go.unpack(gof);
  if (first) {
    c.load(cp + d0); s.load(sp + d0 + half);
    c.unpack(cf); s.unpack(sf);
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = gf[i] * cf[i] + gof[i] * sf[i];
  } else {
    c.load(cp + d0); s.load(sp + d0 - half);
    c.unpack(cf); s.unpack(sf);
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = gf[i] * cf[i] - gof[i] * sf[i];
  }
  Vec16<T> ov;
  ov.pack(o);
  ov.store(dst);
}

template <typename T>
__global__ void quick_gelu_kernel(T* __restrict__ x, int64_t nvec) {
  constexpr int N = Vec16<T>::N;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nvec;
       j += (int64_t)gridDim.x * blockDim.x) {
    Vec16<T> v;
    float f[N];
    v.load(x + j * N);
    v.unpack(f);
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = f[i] / (1.f + __expf(-1.702f * f[i]));
    v.pack(f);
    v.store(x + j * N);
  }
}

// Phi(x) (standard normal CDF) and phi(x) (its density).
template <bool FAST>
__device__ __forceinline__ void normal_cdf_pdf(float x, float& cdf, float& pdf) {
  const float E = __expf(-0.5f * x * x);
  pdf = 0.3989422804014327f * E;
  if (FAST) {
    // erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1/(1 + p z), z = |x|/sqrt2
    // (Abramowitz & Stegun 7.1.26, |err| <= 1.5e-7): shares exp(-x^2/2) with the density.
    const float z = fabsf(x) * 0.7071067811865476f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float half_erfc = 0.5f * p * t * E;          // 0.5 erfc(|z|) = Phi(-|x|)
    cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
  } else {
    cdf = 0.5f * (1.f + erff(x * 0.7071067811865476f));
  }
}

constexpr int GELU_U = 4;   // 16-byte vectors in flight per thread and operand

template <typename T, bool FAST>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                       int64_t nvec) {
  constexpr int N = Vec16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < nvec; j0 += stride * GELU_U) {
    Vec16<T> v[GELU_U];
#pragma unroll
    for (int u = 0; u < GELU_U; ++u)
      if (j0 + u * stride < nvec) v[u].load_stream(x + (j0 + u * stride) * N);
#pragma unroll
    for (int u = 0; u < GELU_U; ++u) {
      if (j0 + u * stride < nvec) {
        float f[N];
        v[u].unpack(f);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          float c, d;
          normal_cdf_pdf<FAST>(f[i], c, d);
          f[i] *= c;
        }
        v[u].pack(f);
        v[u].store(y + (j0 + u * stride) * N);
      }
    }
  }
}

// dx = dy * (Phi(x) + x phi(x))
template <typename T, bool FAST>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                       T* __restrict__ dx, int64_t nvec) {
  constexpr int N = Vec16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < nvec; j0 += stride * GELU_U) {
    Vec16<T> v[GELU_U], g[GELU_U];
#pragma unroll
    for (int u = 0; u < GELU_U; ++u) {
      if (j0 + u * stride < nvec) {
        v[u].load_stream(x + (j0 + u * stride) * N);
        g[u].load_stream(dy + (j0 + u * stride) * N);
      }
    }
#pragma unroll
    for (int u = 0; u < GELU_U; ++u) {
      if (j0 + u * stride < nvec) {
        float f[N], gf[N];
        v[u].unpack(f);
        g[u].unpack(gf);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          float c, d;
          normal_cdf_pdf<FAST>(f[i], c, d);
          gf[i] *= fmaf(f[i], d, c);
        }
        g[u].pack(gf);
        g[u].store(dx + (j0 + u * stride) * N);
      }
    }
  }
}

}  // namespace unimp

using namespace unimp;

static int rotary_check(const char* who, int B, int T, int H, int dh, int rot, int dtype) {
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "%s: dtype", who);
  const int n = dtype == UNIMP_BF16 ? 8 : 4;
  UNIMP_CHECK_ARG(B > 0 && T > 0 && H > 0 && dh > 0 && rot > 0 && rot <= dh && rot % 2 == 0 &&
                      (rot / 2) % n == 0 && dh % n == 0,
                  UNIMP_E_SHAPE, "%s: need rot/2 and dh multiples of %d (dh=%d rot=%d)", who, n, dh, rot);
  return 0;
}

extern "C" int unimp_rotary_qkv_fwd(const void* qkv, void* out, const void* cos, const void* sin, int B,
                                    int T, int H, int dh, int rot, int64_t cs_batch_stride, int dtype,
                                    void* stream) {
  UNIMP_CHECK_ARG(qkv && out && cos && sin && qkv != out, UNIMP_E_NULL,
                  "rotary_qkv_fwd: NULL pointer (or out aliases qkv)");
  int rc = rotary_check("rotary_qkv_fwd", B, T, H, dh, rot, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(aligned16(qkv) && aligned16(out) && aligned16(cos) && aligned16(sin), UNIMP_E_ALIGN,
                  "rotary_qkv_fwd: pointers must be 16-byte aligned");
  const int n = dtype == UNIMP_BF16 ? 8 : 4;
  const int64_t total = (int64_t)B * T * H * 3 * (dh / n);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == UNIMP_BF16)
    rotary_qkv_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, (const __nv_bfloat16*)cos,
        (const __nv_bfloat16*)sin, total, T, H, dh, rot, cs_batch_stride);
  else
    rotary_qkv_fwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float*)qkv, (float*)out, (const float*)cos, (const float*)sin, total, T, H, dh, rot,
        cs_batch_stride);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_rotary_qkv_bwd(const void* dq, const void* dk, const void* dv, const int64_t* strides9,
                                    const void* cos, const void* sin, void* d_qkv, int B, int T, int H,
                                    int dh, int rot, int64_t cs_batch_stride, int dtype, void* stream) {
  UNIMP_CHECK_ARG(dq && dk && dv && strides9 && cos && sin && d_qkv, UNIMP_E_NULL,
                  "rotary_qkv_bwd: NULL pointer");
  int rc = rotary_check("rotary_qkv_bwd", B, T, H, dh, rot, dtype);
  if (rc) return rc;
  const int n = dtype == UNIMP_BF16 ? 8 : 4;
  for (int i = 0; i < 9; ++i)
    UNIMP_CHECK_ARG(strides9[i] % n == 0, UNIMP_E_ALIGN, "rotary_qkv_bwd: stride %d not vector aligned", i);
  UNIMP_CHECK_ARG(aligned16(dq) && aligned16(dk) && aligned16(dv) && aligned16(d_qkv) && aligned16(cos) &&
                      aligned16(sin),
                  UNIMP_E_ALIGN, "rotary_qkv_bwd: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)B * T * H * 3 * (dh / n);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const int64_t* s = strides9;  // HOST array: {q_sb,q_sh,q_st, k_sb,k_sh,k_st, v_sb,v_sh,v_st} in elements
  if (dtype == UNIMP_BF16)
    rotary_qkv_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dq, (const __nv_bfloat16*)dk, (const __nv_bfloat16*)dv, s[0], s[1], s[2], s[3],
        s[4], s[5], s[6], s[7], s[8], (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin,
        (__nv_bfloat16*)d_qkv, total, T, H, dh, rot, cs_batch_stride);
  else
    rotary_qkv_bwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float*)dq, (const float*)dk, (const float*)dv, s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7],
        s[8], (const float*)cos, (const float*)sin, (float*)d_qkv, total, T, H, dh, rot, cs_batch_stride);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_quick_gelu(void* x, int64_t n, int dtype, void* stream) {
  UNIMP_CHECK_ARG(x, UNIMP_E_NULL, "quick_gelu: NULL pointer");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "quick_gelu: dtype");
  const int npv = dtype == UNIMP_BF16 ? 8 : 4;
  UNIMP_CHECK_ARG(n % npv == 0 && aligned16(x), UNIMP_E_ALIGN, "quick_gelu: n %% %d != 0 or unaligned", npv);
  if (n == 0) return 0;
  int64_t blocks = (n / npv + 255) / 256;
  if (blocks > 16 * UNIMP_NUM_SMS) blocks = 16 * UNIMP_NUM_SMS;
  if (dtype == UNIMP_BF16)
    quick_gelu_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, n / npv);
  else
    quick_gelu_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)x, n / npv);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

static int gelu_check(const char* who, const void* a, const void* b, const void* c, int64_t n, int dtype) {
  UNIMP_CHECK_ARG(a && b && c, UNIMP_E_NULL, "%s: NULL pointer", who);
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "%s: dtype", who);
  const int npv = dtype == UNIMP_BF16 ? 8 : 4;
  UNIMP_CHECK_ARG(n >= 0 && n % npv == 0 && aligned16(a) && aligned16(b) && aligned16(c), UNIMP_E_ALIGN,
                  "%s: n %% %d != 0 or unaligned pointer", who, npv);
  return 0;
}

static unsigned gelu_blocks(int64_t nvec) {
  // one pass of GELU_U vectors per thread when it fits; never more than 8 CTAs per SM
  int64_t blocks = (nvec + 256 * GELU_U - 1) / (256 * GELU_U);
  if (blocks > 8 * UNIMP_NUM_SMS) blocks = 8 * UNIMP_NUM_SMS;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

extern "C" int unimp_gelu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream) {
  int rc = gelu_check("gelu_fwd", x, y, y, n, dtype);
  if (rc) return rc;
  if (n == 0) return 0;
  const int64_t nvec = n / (dtype == UNIMP_BF16 ? 8 : 4);
  if (dtype == UNIMP_BF16)
    gelu_fwd_kernel<__nv_bfloat16, true><<<gelu_blocks(nvec), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec);
  else
    gelu_fwd_kernel<float, false><<<gelu_blocks(nvec), 256, 0, (cudaStream_t)stream>>>(
        (const float*)x, (float*)y, nvec);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_gelu_bwd(const void* x, const void* dy, void* dx, int64_t n, int dtype, void* stream) {
  int rc = gelu_check("gelu_bwd", x, dy, dx, n, dtype);
  if (rc) return rc;
  if (n == 0) return 0;
  const int64_t nvec = n / (dtype == UNIMP_BF16 ? 8 : 4);
  if (dtype == UNIMP_BF16)
    gelu_bwd_kernel<__nv_bfloat16, true><<<gelu_blocks(nvec), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, nvec);
  else
    gelu_bwd_kernel<float, false><<<gelu_blocks(nvec), 256, 0, (cudaStream_t)stream>>>(
        (const float*)x, (const float*)dy, (float*)dx, nvec);
  UNIMP_CHECK_LAUNCH();
  return 0;
}
