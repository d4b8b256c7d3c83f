// fp32-exact attention cores on CUDA cores (any storage dtype, fp32 math).
//
// This is the path used when the model runs in fp32 (BASELINE.json: "fp32 rel 1e-4" parity
// needs full-precision products, which the bf16 tensor-core path cannot give), the backward of
// the media-located mask's odd corners, and the decode step.  The bf16 hot path lives in
// attn_tc.cu (tcgen05 + TMEM + TMA).
//
// Masking (MaskedCrossAttention): query row i of sample b attends keys
//   [ (tt-1)*n, tt*n )  with tt = text_time[b,i];  tt == 0 -> nothing (output 0);
//   tt > Ti -> every key with a CONSTANT score (upstream's all-masked row: softmax of a
//   constant is uniform, and that constant does not depend on q or k => dq = dk = 0).
#include "common.cuh"

namespace unimp {

constexpr int SQ = 16;   // query rows per tile
constexpr int SK = 64;   // keys per block
constexpr int SD = 64;   // head dim
constexpr int SP = SD + 1;  // padded row stride (floats): conflict-free column walks
constexpr int SIMT_THREADS = 256;

struct RowKeys {
  int lo, hi;
  bool uniform;
};

__device__ __forceinline__ RowKeys row_keys(const int32_t* text_time, int b, int i, int Lq, int Lk,
                                            int n, int Ti) {
  RowKeys r;
  if (!text_time) {
    r.lo = 0; r.hi = Lk; r.uniform = false;
    return r;
  }
  const int tt = text_time[(int64_t)b * Lq + i];
  if (tt <= 0) { r.lo = 0; r.hi = 0; r.uniform = false; }
  else if (tt > Ti) { r.lo = 0; r.hi = Lk; r.uniform = true; }
  else { r.lo = (tt - 1) * n; r.hi = tt * n; r.uniform = false; }
  return r;
}

template <typename T>
__device__ __forceinline__ void stage_rows(float* dst, int dst_stride, const T* base,
                                           int64_t row_stride, int r0, int nrows, int rmax,
                                           float mul) {
  // dst[r][d] = mul * src[r0 + r][d], zero beyond rmax; 64 columns.
  for (int idx = threadIdx.x; idx < nrows * (SD / 4); idx += blockDim.x) {
    const int r = idx / (SD / 4), c = (idx % (SD / 4)) * 4;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0 + r < rmax) {
      const T* p = base + (int64_t)(r0 + r) * row_stride + c;
#pragma unroll
      for (int e = 0; e < 4; ++e) f[e] = Elem<T>::to_f(p[e]) * mul;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[r * dst_stride + c + e] = f[e];
  }
}

// grid (ceil(Lq/SQ), H, B)
template <typename T>
__global__ void __launch_bounds__(SIMT_THREADS)
attn_fwd_simt_kernel(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* __restrict__ tt,
                     unimp_mview_t o, float* __restrict__ lse, int Lq, int Lk, int H, int n, int Ti,
                     float scale) {
  __shared__ float sQ[SQ * SP], sK[SK * SP], sV[SK * SP], sS[SQ * SK];
  __shared__ float sM[SQ], sL[SQ], sC[SQ];
  __shared__ int sLo[SQ], sHi[SQ], sUni[SQ];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * SQ;
  const T* qb = (const T*)q.ptr + b * q.batch_stride + h * SD;
  const T* kb = (const T*)k.ptr + b * k.batch_stride + h * SD;
  const T* vb = (const T*)v.ptr + b * v.batch_stride + h * SD;
  stage_rows<T>(sQ, SP, qb, q.row_stride, q0, SQ, Lq, scale);
  if (threadIdx.x < SQ) {
    const int i = q0 + threadIdx.x;
    RowKeys rk = {0, 0, false};
    if (i < Lq) rk = row_keys(tt, b, i, Lq, Lk, n, Ti);
    sLo[threadIdx.x] = rk.lo; sHi[threadIdx.x] = rk.hi; sUni[threadIdx.x] = rk.uniform;
    sM[threadIdx.x] = -INFINITY; sL[threadIdx.x] = 0.f;
  }
  __syncthreads();
  int lo = Lk, hi = 0;
  for (int r = 0; r < SQ; ++r) { lo = min(lo, sLo[r]); hi = max(hi, sHi[r]); }
  const int oi = threadIdx.x / 16, oc = (threadIdx.x % 16) * 4;  // output row / 4 columns
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = (lo / SK) * SK; k0 < hi; k0 += SK) {
    __syncthreads();
    stage_rows<T>(sK, SP, kb, k.row_stride, k0, SK, Lk, 1.f);
    stage_rows<T>(sV, SP, vb, v.row_stride, k0, SK, Lk, 1.f);
    __syncthreads();
    // scores: thread -> key j = tid%64, rows i = tid/64 + 4r
    {
      const int j = threadIdx.x % SK;
#pragma unroll
      for (int r = 0; r < SQ / 4; ++r) {
        const int i = threadIdx.x / SK + 4 * r;
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < SD; ++d) s = fmaf(sQ[i * SP + d], sK[j * SP + d], s);
        const int kj = k0 + j;
        if (sUni[i]) s = 0.f;
        if (kj < sLo[i] || kj >= sHi[i]) s = -INFINITY;
        sS[i * SK + j] = s;
      }
    }
    __syncthreads();
    // online softmax per row: warp w owns rows 2w, 2w+1
    {
      const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int i = 2 * w + rr;
        const float s0 = sS[i * SK + lane], s1 = sS[i * SK + lane + 32];
        const float mo = sM[i];
        const float mn = fmaxf(mo, warp_max(fmaxf(s0, s1)));
        float p0 = 0.f, p1 = 0.f, corr = 1.f;
        if (mn > -INFINITY) {
          p0 = __expf(s0 - mn); p1 = __expf(s1 - mn);
          corr = __expf(mo - mn);  // mo == -inf -> 0
        }
        const float ps = warp_sum(p0 + p1);
        sS[i * SK + lane] = p0; sS[i * SK + lane + 32] = p1;
        if (lane == 0) { sM[i] = mn; sL[i] = sL[i] * corr + ps; sC[i] = corr; }
      }
    }
    __syncthreads();
    {
      const float c = sC[oi];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] *= c;
      for (int j = 0; j < SK; ++j) {
        const float p = sS[oi * SK + j];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(p, sV[j * SP + oc + e], acc[e]);
      }
    }
  }
  __syncthreads();
  const int i = q0 + oi;
  if (i < Lq) {
    const float l = sL[oi];
    const float inv = l > 0.f ? 1.f / l : 0.f;
    T* op = (T*)o.ptr + b * o.batch_stride + (int64_t)i * o.row_stride + h * SD + oc;
#pragma unroll
    for (int e = 0; e < 4; ++e) op[e] = Elem<T>::from_f(acc[e] * inv);
    if (oc == 0) lse[((int64_t)b * H + h) * Lq + i] = l > 0.f ? sM[oi] + logf(l) : -INFINITY;
  }
}

// delta[b,h,i] = sum_d dO*O ; also zeroes nothing else. grid (ceil(Lq/8), H, B), 256 threads
template <typename T>
__global__ void attn_delta_kernel(unimp_view_t o, unimp_view_t d_o, float* __restrict__ delta, int Lq,
                                  int H) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= Lq) return;
  const T* op = (const T*)o.ptr + b * o.batch_stride + (int64_t)i * o.row_stride + h * SD;
  const T* gp = (const T*)d_o.ptr + b * d_o.batch_stride + (int64_t)i * d_o.row_stride + h * SD;
  float s = Elem<T>::to_f(op[lane]) * Elem<T>::to_f(gp[lane]) +
            Elem<T>::to_f(op[lane + 32]) * Elem<T>::to_f(gp[lane + 32]);
  s = warp_sum(s);
  if (lane == 0) delta[((int64_t)b * H + h) * Lq + i] = s;
}

// One CTA per (key block, h, b); loops over query tiles; dK/dV in registers;
// dQ accumulated with fp32 atomics into dq_acc (B, Lq, H, 64) (single writer when masked).
template <typename T>
__global__ void __launch_bounds__(SIMT_THREADS)
attn_bwd_simt_kernel(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* __restrict__ tt,
                     unimp_view_t d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                     float* __restrict__ dq_acc, unimp_mview_t dk, unimp_mview_t dv, int Lq, int Lk,
                     int H, int n, int Ti, float scale) {
  extern __shared__ float smem[];
  float* sQ = smem;                 // SQ x SP (pre-scaled)
  float* sG = sQ + SQ * SP;         // dO  SQ x SP
  float* sK = sG + SQ * SP;         // SK x SP
  float* sV = sK + SK * SP;         // SK x SP
  float* sP = sV + SK * SP;         // SQ x SK
  float* sDS = sP + SQ * SK;        // SQ x SK
  __shared__ float sLse[SQ], sDel[SQ];
  __shared__ int sLo[SQ], sHi[SQ], sUni[SQ];
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * SK;
  const T* qb = (const T*)q.ptr + b * q.batch_stride + h * SD;
  const T* gb = (const T*)d_o.ptr + b * d_o.batch_stride + h * SD;
  const T* kb = (const T*)k.ptr + b * k.batch_stride + h * SD;
  const T* vb = (const T*)v.ptr + b * v.batch_stride + h * SD;
  stage_rows<T>(sK, SP, kb, k.row_stride, k0, SK, Lk, 1.f);
  stage_rows<T>(sV, SP, vb, v.row_stride, k0, SK, Lk, 1.f);
  const int kj = threadIdx.x % SK;        // key owned for dK/dV
  const int dc = (threadIdx.x / SK) * 16; // 16 d-columns owned for dK/dV
  float aK[16], aV[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) aK[e] = aV[e] = 0.f;
  for (int q0 = 0; q0 < Lq; q0 += SQ) {
    __syncthreads();
    if (threadIdx.x < SQ) {
      const int i = q0 + threadIdx.x;
      RowKeys rk = {0, 0, false};
      float l = 0.f, dl = 0.f;
      if (i < Lq) {
        rk = row_keys(tt, b, i, Lq, Lk, n, Ti);
        l = lse[((int64_t)b * H + h) * Lq + i];
        dl = delta[((int64_t)b * H + h) * Lq + i];
      }
      sLo[threadIdx.x] = rk.lo; sHi[threadIdx.x] = rk.hi; sUni[threadIdx.x] = rk.uniform;
      sLse[threadIdx.x] = l; sDel[threadIdx.x] = dl;
    }
    __syncthreads();
    bool any = false;
    for (int r = 0; r < SQ; ++r) any |= (sLo[r] < k0 + SK && sHi[r] > k0);
    if (!any) continue;  // uniform across the CTA
    stage_rows<T>(sQ, SP, qb, q.row_stride, q0, SQ, Lq, scale);
    stage_rows<T>(sG, SP, gb, d_o.row_stride, q0, SQ, Lq, 1.f);
    __syncthreads();
    {
      const int j = threadIdx.x % SK;
#pragma unroll
      for (int r = 0; r < SQ / 4; ++r) {
        const int i = threadIdx.x / SK + 4 * r;
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < SD; ++d) {
          s = fmaf(sQ[i * SP + d], sK[j * SP + d], s);
          dp = fmaf(sG[i * SP + d], sV[j * SP + d], dp);
        }
        const int kk = k0 + j;
        float p = 0.f, ds = 0.f;
        if (kk >= sLo[i] && kk < sHi[i] && kk < Lk) {
          if (sUni[i]) { p = __expf(-sLse[i]); ds = 0.f; }
          else { p = __expf(s - sLse[i]); ds = p * (dp - sDel[i]); }
        }
        sP[i * SK + j] = p;
        sDS[i * SK + j] = ds;
      }
    }
    __syncthreads();
    // dV[kj][dc..] += P[i][kj] dO[i][dc..] ; dK[kj][dc..] += dS[i][kj] Qs[i][dc..]
    for (int i = 0; i < SQ; ++i) {
      const float p = sP[i * SK + kj], ds = sDS[i * SK + kj];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        aV[e] = fmaf(p, sG[i * SP + dc + e], aV[e]);
        aK[e] = fmaf(ds, sQ[i * SP + dc + e], aK[e]);  // sQ is pre-scaled => includes `scale`
      }
    }
    // dQ[i][c..c+3] = scale * sum_j dS[i][j] K[j][c..]
    {
      const int i = threadIdx.x / 16, c = (threadIdx.x % 16) * 4;
      if (q0 + i < Lq && sLo[i] < k0 + SK && sHi[i] > k0 && !sUni[i]) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < SK; ++j) {
          const float ds = sDS[i * SK + j];
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] = fmaf(ds, sK[j * SP + c + e], a[e]);
        }
        float* dst = dq_acc + (((int64_t)b * Lq + q0 + i) * H + h) * SD + c;
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(dst + e, a[e] * scale);
      }
    }
  }
  if (k0 + kj < Lk) {
    T* dkp = (T*)dk.ptr + b * dk.batch_stride + (int64_t)(k0 + kj) * dk.row_stride + h * SD + dc;
    T* dvp = (T*)dv.ptr + b * dv.batch_stride + (int64_t)(k0 + kj) * dv.row_stride + h * SD + dc;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      dkp[e] = Elem<T>::from_f(aK[e]);
      dvp[e] = Elem<T>::from_f(aV[e]);
    }
  }
}

template <typename T>
__global__ void dq_convert_kernel(const float* __restrict__ acc, unimp_mview_t dq, int Lq, int H) {
  // one thread per 4 elements of (b, i, h, d)
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_b = (int64_t)Lq * H * (SD / 4);
  const int b = blockIdx.y;
  if (idx >= per_b) return;
  const int c = (int)(idx % (SD / 4)) * 4;
  const int h = (int)((idx / (SD / 4)) % H);
  const int i = (int)(idx / ((SD / 4) * H));
  const float* src = acc + (((int64_t)b * Lq + i) * H + h) * SD + c;
  T* dst = (T*)dq.ptr + b * dq.batch_stride + (int64_t)i * dq.row_stride + h * SD + c;
#pragma unroll
  for (int e = 0; e < 4; ++e) dst[e] = Elem<T>::from_f(src[e]);
}

// Decode: one CTA per (b, h); the single query attends the n keys of image n_media[b]-1 (all Ti*n
// keys, uniformly, if n_media[b] > Ti).  A thread owns one key — its whole 64-wide dot product,
// independent 16-byte loads — the softmax is a block reduction, and for P.V the 256 threads are 4 key
// groups x 64 output columns (coalesced V rows, 16 independent loads each), folded in shared memory.
// (Round 2's first version walked the keys serially in one warp: 64 dependent load -> shuffle -> exp
// rounds, 38 us per launch, 0.6 ms of a 6.3 ms decode step.)
constexpr int XD_THREADS = 256;
template <typename T>
__global__ void __launch_bounds__(XD_THREADS) xattn_decode_kernel(unimp_view_t q, unimp_view_t k, unimp_view_t v,
                                                                  const int32_t* __restrict__ n_media,
                                                                  unimp_mview_t o, int Ti, int n, int H,
                                                                  float scale) {
  extern __shared__ float xd_p[];           // one probability per attended key
  __shared__ float sq[SD], sred[32], sacc[XD_THREADS];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, h = blockIdx.x, tid = threadIdx.x;
  const int tt = n_media[b];
  T* op = (T*)o.ptr + b * o.batch_stride + h * SD;
  if (tt <= 0) {
    if (tid < SD) op[tid] = Elem<T>::from_f(0.f);
    return;
  }
  const bool uni = tt > Ti;
  const int lo = uni ? 0 : (tt - 1) * n, nk = uni ? Ti * n : n;
  const T* qp = (const T*)q.ptr + b * q.batch_stride + h * SD;
  if (tid < SD) sq[tid] = Elem<T>::to_f(qp[tid]) * scale;
  __syncthreads();
  float m = -INFINITY;
  for (int j = tid; j < nk; j += XD_THREADS) {
    float s = 0.f;
    if (!uni) {
      const T* kp = (const T*)k.ptr + b * k.batch_stride + (int64_t)(lo + j) * k.row_stride + h * SD;
      if ((reinterpret_cast<uintptr_t>(kp) & 15u) == 0) {
        constexpr int N = Vec16<T>::N;
        Vec16<T> kv[8];
#pragma unroll
        for (int c0 = 0; c0 < SD / N; c0 += 8) {
#pragma unroll
          for (int c = 0; c < 8; ++c) kv[c].load(kp + (c0 + c) * N);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float f[N];
            kv[c].unpack(f);
#pragma unroll
            for (int i = 0; i < N; ++i) s = fmaf(sq[(c0 + c) * N + i], f[i], s);
          }
        }
      } else {
#pragma unroll 16
        for (int d = 0; d < SD; ++d) s = fmaf(sq[d], Elem<T>::to_f(kp[d]), s);
      }
    }
    xd_p[j] = s;
    m = fmaxf(m, s);
  }
  m = block_max(m, sred);
  float l = 0.f;
  for (int j = tid; j < nk; j += XD_THREADS) {
    const float p = __expf(xd_p[j] - m);
    xd_p[j] = p;
    l += p;
  }
  l = block_sum(l, sred);                   // (its barriers also publish xd_p)
  const int g = tid >> 6, d = tid & (SD - 1);
  const T* vp = (const T*)v.ptr + b * v.batch_stride + (int64_t)lo * v.row_stride + h * SD + d;
  float acc = 0.f;
#pragma unroll 8
  for (int j = g; j < nk; j += XD_THREADS / SD) acc = fmaf(xd_p[j], Elem<T>::to_f(vp[(int64_t)j * v.row_stride]), acc);
  sacc[tid] = acc;
  __syncthreads();
  if (tid < SD) op[tid] = Elem<T>::from_f((sacc[tid] + sacc[tid + 64] + sacc[tid + 128] + sacc[tid + 192]) / l);
}

// ---- host launchers (used by capi.cu) -----------------------------------------------------

template <typename T>
int launch_attn_fwd_simt(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                         unimp_mview_t o, float* lse, int B, int Lq, int Lk, int H, int n, int Ti,
                         float scale, cudaStream_t st) {
  dim3 grid((Lq + SQ - 1) / SQ, H, B);
  attn_fwd_simt_kernel<T><<<grid, SIMT_THREADS, 0, st>>>(q, k, v, tt, o, lse, Lq, Lk, H, n, Ti, scale);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int launch_attn_bwd_simt(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                         unimp_view_t o, unimp_view_t d_o, const float* lse, void* workspace,
                         unimp_mview_t dq, unimp_mview_t dk, unimp_mview_t dv, int B, int Lq, int Lk,
                         int H, int n, int Ti, float scale, cudaStream_t st) {
  float* delta = (float*)workspace;
  float* dq_acc = delta + (((int64_t)B * H * Lq + 3) / 4) * 4;
  const size_t dq_bytes = (size_t)B * Lq * H * SD * sizeof(float);
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, dq_bytes, st);
  if (e != cudaSuccess) { set_error("attn_bwd: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  attn_delta_kernel<T><<<dim3((Lq + 7) / 8, H, B), 256, 0, st>>>(o, d_o, delta, Lq, H);
  UNIMP_CHECK_LAUNCH();
  const int smem = (2 * SQ * SP + 2 * SK * SP + 2 * SQ * SK) * sizeof(float);
  static bool attr_set[2] = {false, false};
  const int ti = sizeof(T) == 4 ? 0 : 1;
  if (!attr_set[ti]) {
    cudaFuncSetAttribute(attn_bwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set[ti] = true;
  }
  dim3 grid((Lk + SK - 1) / SK, H, B);
  attn_bwd_simt_kernel<T><<<grid, SIMT_THREADS, smem, st>>>(q, k, v, tt, d_o, lse, delta, dq_acc, dk,
                                                            dv, Lq, Lk, H, n, Ti, scale);
  UNIMP_CHECK_LAUNCH();
  const int64_t per_b = (int64_t)Lq * H * (SD / 4);
  dq_convert_kernel<T><<<dim3((unsigned)((per_b + 255) / 256), B), 256, 0, st>>>(dq_acc, dq, Lq, H);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int launch_xattn_decode(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* n_media,
                        unimp_mview_t o, int B, int Ti, int n, int H, float scale, cudaStream_t st) {
  cudaError_t e = launch_pdl(xattn_decode_kernel<T>, dim3(H, B), dim3(XD_THREADS), (size_t)Ti * n * sizeof(float), st,
                             q, k, v, n_media, o, Ti, n, H, scale);
  if (e != cudaSuccess) { set_error("xattn_decode launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

#define INST(T)                                                                                     \
  template int launch_attn_fwd_simt<T>(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*,    \
                                       unimp_mview_t, float*, int, int, int, int, int, int, float,  \
                                       cudaStream_t);                                               \
  template int launch_attn_bwd_simt<T>(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*,    \
                                       unimp_view_t, unimp_view_t, const float*, void*,             \
                                       unimp_mview_t, unimp_mview_t, unimp_mview_t, int, int, int,  \
                                       int, int, int, float, cudaStream_t);                         \
  template int launch_xattn_decode<T>(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*,     \
                                      unimp_mview_t, int, int, int, int, float, cudaStream_t);
INST(float)
INST(__nv_bfloat16)

}  // namespace unimp
