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

// Work decomposition shared by the two rotary kernels.  One CTA per (b, t) row of the packed
// (B, T, H, 3, dh) projection; a work item is (head, {q,k,v}, slot): slot p < hv owns the vector
// pair (p, p + hv) of the two rotary halves (hv = rot/2 / N) — the rotation reads x1 and x2 once
// and writes both outputs, instead of every output vector re-reading its partner — and slots
// p >= hv copy one vector of the non-rotary tail.  32-bit index math only (the previous
// one-vector-per-thread version spent its time in six 64-bit divisions per thread).
struct RotaryGeom {
  int hv, P, items;   // vectors per rotary half, slots per (head, part), items per row
};

__device__ __forceinline__ void rotary_item(int it, const RotaryGeom& g, int& h, int& which, int& p) {
  const unsigned hw = (unsigned)it / (unsigned)g.P;
  p = it - (int)hw * g.P;
  h = (int)(hw / 3u);
  which = (int)(hw - 3u * (unsigned)h);
}

// out[b,t,h,{q,k,v},:] = rotary(qkv) (v and the non-rotary tail copied).
// cos/sin: (cb, T, rot), HF layout (the two halves need not be duplicates: both are read).
template <typename T>
__global__ void __launch_bounds__(256) rotary_qkv_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                             const T* __restrict__ cs, const T* __restrict__ sn,
                                                             int Tn, int H, int dh, int rot, int64_t cs_bstride,
                                                             RotaryGeom g) {
  constexpr int N = Vec16<T>::N;
  const int64_t row = blockIdx.x;
  const int t = (int)(row % Tn);
  const int64_t b = row / Tn;
  const int half = rot / 2;
  const T* cp = cs + b * cs_bstride + (int64_t)t * rot;
  const T* sp = sn + b * cs_bstride + (int64_t)t * rot;
  const T* srow = qkv + row * H * 3 * dh;
  T* drow = out + row * H * 3 * dh;
  for (int it = threadIdx.x; it < g.items; it += blockDim.x) {
    int h, which, p;
    rotary_item(it, g, h, which, p);
    const int off = (h * 3 + which) * dh;
    if (p >= g.hv) {                               // non-rotary tail: one vector
      const int d = rot + (p - g.hv) * N;
      Vec16<T> x;
      x.load_stream(srow + off + d);
      x.store(drow + off + d);
      continue;
    }
    const int d0 = p * N;
    Vec16<T> x1, x2;
    x1.load_stream(srow + off + d0);
    x2.load_stream(srow + off + d0 + half);
    if (which == 2) {
      x1.store(drow + off + d0);
      x2.store(drow + off + d0 + half);
      continue;
    }
    // q*cos + rotate_half(q)*sin: first half: x1 c1 - x2 s1 ; second half: x2 c2 + x1 s2
    Vec16<T> c1, s1, c2, s2;
    c1.load(cp + d0); s1.load(sp + d0);
    c2.load(cp + d0 + half); s2.load(sp + d0 + half);
    float a[N], bb[N], c1f[N], s1f[N], c2f[N], s2f[N], o1[N], o2[N];
    x1.unpack(a); x2.unpack(bb); c1.unpack(c1f); s1.unpack(s1f); c2.unpack(c2f); s2.unpack(s2f);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      o1[i] = a[i] * c1f[i] - bb[i] * s1f[i];
      o2[i] = bb[i] * c2f[i] + a[i] * s2f[i];
    }
    x1.pack(o1); x2.pack(o2);
    x1.store(drow + off + d0);
    x2.store(drow + off + d0 + half);
  }
}

// d_qkv[b,t,h,{q,k,v},:] from dq/dk/dv (B,H,T,dh) strided views: un-rotates dq/dk and packs
// dq|dk|dv in one pass.
// N consecutive gradient elements -> fp32 (S = the element type of the source: T, or float for the
// fp32 dq scratch of unimp_lm_attn_bwd)
template <typename S, int N>
__device__ __forceinline__ void load_grad(const S* p, float* f) {
  if constexpr (sizeof(S) == 4 && N == 8) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    Vec16<S> v;
    v.load_stream(p);
    v.unpack(f);
  }
}

template <typename T, typename QS>
__global__ void __launch_bounds__(256) rotary_qkv_bwd_kernel(
    const QS* __restrict__ dq, const T* __restrict__ dk, const T* __restrict__ dv, int64_t q_sb, int64_t q_sh,
    int64_t q_st, int64_t k_sb, int64_t k_sh, int64_t k_st, int64_t v_sb, int64_t v_sh, int64_t v_st,
    const T* __restrict__ cs, const T* __restrict__ sn, T* __restrict__ d_qkv, int Tn, int H, int dh, int rot,
    int64_t cs_bstride, RotaryGeom g) {
  constexpr int N = Vec16<T>::N;
  const int64_t row = blockIdx.x;
  const int t = (int)(row % Tn);
  const int64_t b = row / Tn;
  const int half = rot / 2;
  const T* cp = cs + b * cs_bstride + (int64_t)t * rot;
  const T* sp = sn + b * cs_bstride + (int64_t)t * rot;
  const QS* qb = dq + b * q_sb + (int64_t)t * q_st;
  const T* kb = dk + b * k_sb + (int64_t)t * k_st;
  const T* vb = dv + b * v_sb + (int64_t)t * v_st;
  T* drow = d_qkv + row * H * 3 * dh;
  for (int it = threadIdx.x; it < g.items; it += blockDim.x) {
    int h, which, p;
    rotary_item(it, g, h, which, p);
    const T* src = which == 1 ? kb + h * k_sh : vb + h * v_sh;
    const QS* qsrc = qb + h * q_sh;
    T* dst = drow + (h * 3 + which) * dh;
    float a[N], bb[N];
    Vec16<T> g1, g2;
    if (p >= g.hv) {
      const int d = rot + (p - g.hv) * N;
      if (which == 0) load_grad<QS, N>(qsrc + d, a); else load_grad<T, N>(src + d, a);
      g1.pack(a);
      g1.store(dst + d);
      continue;
    }
    const int d0 = p * N;
    if (which == 0) {
      load_grad<QS, N>(qsrc + d0, a);
      load_grad<QS, N>(qsrc + d0 + half, bb);
    } else {
      load_grad<T, N>(src + d0, a);
      load_grad<T, N>(src + d0 + half, bb);
    }
    if (which == 2) {
      g1.pack(a); g2.pack(bb);
      g1.store(dst + d0);
      g2.store(dst + d0 + half);
      continue;
    }
    // out1 = x1 c1 - x2 s1 ; out2 = x2 c2 + x1 s2   =>   dx1 = g1 c1 + g2 s2 ; dx2 = g2 c2 - g1 s1
    Vec16<T> c1, s1, c2, s2;
    c1.load(cp + d0); s1.load(sp + d0);
    c2.load(cp + d0 + half); s2.load(sp + d0 + half);
    float c1f[N], s1f[N], c2f[N], s2f[N], o1[N], o2[N];
    c1.unpack(c1f); s1.unpack(s1f); c2.unpack(c2f); s2.unpack(s2f);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      o1[i] = a[i] * c1f[i] + bb[i] * s2f[i];
      o2[i] = bb[i] * c2f[i] - a[i] * s1f[i];
    }
    g1.pack(o1); g2.pack(o2);
    g1.store(dst + d0);
    g2.store(dst + d0 + half);
  }
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

// Phi(-|x|) (the small tail of the standard normal CDF) and E = exp(-x^2/2).
//   FAST: erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1/(1 + p z), z = |x|/sqrt2
//   (Abramowitz & Stegun 7.1.26, |err| <= 1.5e-7); it shares exp(-x^2/2) with the density, the
//   1/2 and the 1/sqrt2 are folded into the constants: ~15 instructions per element with 2 MUFU.
template <bool FAST>
__device__ __forceinline__ float normal_tail(float x, float& E) {
  const float ax = fabsf(x);
  E = __expf(-0.5f * x * x);
  if (FAST) {
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.7071067811865476f, ax, 1.f)));
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    return p * t * E;                                   // 0.5 erfc(|x|/sqrt2)
  }
  return 0.5f * erfcf(ax * 0.7071067811865476f);
}

// gelu(x) = x Phi(x) = max(x, 0) - |x| Phi(-|x|)
template <bool FAST>
__device__ __forceinline__ float gelu_value(float x) {
  float E;
  const float h = normal_tail<FAST>(x, E);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}

// d gelu / dx = Phi(x) + x phi(x)
template <bool FAST>
__device__ __forceinline__ float gelu_slope(float x) {
  float E;
  const float h = normal_tail<FAST>(x, E);
  const float cdf = x >= 0.f ? 1.f - h : h;
  return fmaf(x * 0.3989422804014327f, E, cdf);
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
        for (int i = 0; i < N; ++i) f[i] = gelu_value<FAST>(f[i]);
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
        for (int i = 0; i < N; ++i) gf[i] *= gelu_slope<FAST>(f[i]);
        g[u].pack(gf);
        g[u].store(dx + (j0 + u * stride) * N);
      }
    }
  }
}

}  // namespace unimp

using namespace unimp;

static RotaryGeom rotary_geom(int H, int dh, int rot, int dtype) {
  const int n = dtype == UNIMP_BF16 ? 8 : 4;
  RotaryGeom g;
  g.hv = (rot / 2) / n;
  g.P = g.hv + (dh - rot) / n;
  g.items = H * 3 * g.P;
  return g;
}

static int rotary_threads(const RotaryGeom& g) {
  // whole passes over the row's items: 480 items (4B: 32 heads x 3 x 5 slots) -> 160 threads x 3
  int best = 256;
  for (int th = 256; th >= 96; th -= 32)
    if (g.items % th == 0) { best = th; break; }
  if (g.items < best) best = ((g.items + 31) / 32) * 32;
  return best;
}

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
  const RotaryGeom g = rotary_geom(H, dh, rot, dtype);
  const unsigned rows = (unsigned)((int64_t)B * T);
  const int threads = rotary_threads(g);
  if (dtype == UNIMP_BF16)
    rotary_qkv_fwd_kernel<__nv_bfloat16><<<rows, threads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, (const __nv_bfloat16*)cos,
        (const __nv_bfloat16*)sin, T, H, dh, rot, cs_batch_stride, g);
  else
    rotary_qkv_fwd_kernel<float><<<rows, threads, 0, (cudaStream_t)stream>>>(
        (const float*)qkv, (float*)out, (const float*)cos, (const float*)sin, T, H, dh, rot,
        cs_batch_stride, g);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

static int rotary_bwd_impl(const char* who, const void* dq, int dq_f32, const void* dk, const void* dv,
                           const int64_t* strides9, const void* cos, const void* sin, void* d_qkv, int B, int T,
                           int H, int dh, int rot, int64_t cs_batch_stride, int dtype, void* stream) {
  UNIMP_CHECK_ARG(dq && dk && dv && strides9 && cos && sin && d_qkv, UNIMP_E_NULL, "%s: NULL pointer", who);
  int rc = rotary_check(who, B, T, H, dh, rot, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(!dq_f32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "%s: an fp32 dq needs dtype bf16", who);
  const int n = dtype == UNIMP_BF16 ? 8 : 4;
  for (int i = 0; i < 9; ++i)
    UNIMP_CHECK_ARG(strides9[i] % (i < 3 && dq_f32 ? 4 : n) == 0, UNIMP_E_ALIGN, "%s: stride %d not vector aligned",
                    who, i);
  UNIMP_CHECK_ARG(aligned16(dq) && aligned16(dk) && aligned16(dv) && aligned16(d_qkv) && aligned16(cos) &&
                      aligned16(sin),
                  UNIMP_E_ALIGN, "%s: pointers must be 16-byte aligned", who);
  const RotaryGeom g = rotary_geom(H, dh, rot, dtype);
  const unsigned rows = (unsigned)((int64_t)B * T);
  const int threads = rotary_threads(g);
  const int64_t* s = strides9;  // HOST array: {q_sb,q_sh,q_st, k_sb,k_sh,k_st, v_sb,v_sh,v_st} in elements
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNIMP_BF16 && dq_f32)
    rotary_qkv_bwd_kernel<__nv_bfloat16, float><<<rows, threads, 0, st>>>(
        (const float*)dq, (const __nv_bfloat16*)dk, (const __nv_bfloat16*)dv, s[0], s[1], s[2], s[3], s[4], s[5],
        s[6], s[7], s[8], (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin, (__nv_bfloat16*)d_qkv, T, H, dh,
        rot, cs_batch_stride, g);
  else if (dtype == UNIMP_BF16)
    rotary_qkv_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<rows, threads, 0, st>>>(
        (const __nv_bfloat16*)dq, (const __nv_bfloat16*)dk, (const __nv_bfloat16*)dv, s[0], s[1], s[2], s[3],
        s[4], s[5], s[6], s[7], s[8], (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin,
        (__nv_bfloat16*)d_qkv, T, H, dh, rot, cs_batch_stride, g);
  else
    rotary_qkv_bwd_kernel<float, float><<<rows, threads, 0, st>>>(
        (const float*)dq, (const float*)dk, (const float*)dv, s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7],
        s[8], (const float*)cos, (const float*)sin, (float*)d_qkv, T, H, dh, rot, cs_batch_stride, g);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_rotary_qkv_bwd(const void* dq, const void* dk, const void* dv, const int64_t* strides9,
                                    const void* cos, const void* sin, void* d_qkv, int B, int T, int H,
                                    int dh, int rot, int64_t cs_batch_stride, int dtype, void* stream) {
  return rotary_bwd_impl("rotary_qkv_bwd", dq, 0, dk, dv, strides9, cos, sin, d_qkv, B, T, H, dh, rot,
                         cs_batch_stride, dtype, stream);
}

extern "C" int unimp_rotary_qkv_bwd_f32q(const float* dq, const void* dk, const void* dv, const int64_t* strides9,
                                         const void* cos, const void* sin, void* d_qkv, int B, int T, int H,
                                         int dh, int rot, int64_t cs_batch_stride, int dtype, void* stream) {
  return rotary_bwd_impl("rotary_qkv_bwd_f32q", dq, 1, dk, dv, strides9, cos, sin, d_qkv, B, T, H, dh, rot,
                         cs_batch_stride, dtype, stream);
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
