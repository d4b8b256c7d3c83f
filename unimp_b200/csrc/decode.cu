// Decode-side kernels of the language model (SURVEY.md §8 a12 / f4, configs[3]: explanation generation,
// reference call `UniMP/pipeline/eval/eval_exp.py:101-114` -> upstream `Flamingo.generate` -> HF beam
// search).  One new token per beam and step: everything here is HBM-bound streaming work.
//
//   lm_decode_attn_kernel   the GPT-NeoX self-attention of ONE new token against static K/V caches:
//                           rotary on the packed qkv projection, the new key/value written at the
//                           device-side cursor, scores / softmax / PV over the cached positions.  Beam
//                           re-ordering is an INDIRECTION table (cache row per (beam, position)) read by
//                           the kernel, not a copy of the caches: HF's `reorder_cache` (and round 2's first
//                           decoder) moved 2 x 32 x 20 MB per token, 1.4 ms of a 6.3 ms step.
//   linear_small_m_kernel   y = act(x W^T + b) for <= 8 rows of x (the beams): the weight matrix is
//                           streamed once with 16-byte loads straight into `mma.sync` fragments (the
//                           legacy warp-level tensor path: the FLOPs are irrelevant, the point is that
//                           the SIMT pipes do not have to unpack and multiply 5 x 3.6 G weights per
//                           token), split-K over the 8 warps of a CTA, fixed-order shared-memory fold.
#include <stdlib.h>

#include "common.cuh"

namespace unimp {

// ------------------------------------------------------------------------------------------------
// K4-decode
// ------------------------------------------------------------------------------------------------
constexpr int LD_THREADS = 256;

template <typename T, int DH>
__global__ void __launch_bounds__(LD_THREADS)
lm_decode_attn_kernel(const T* __restrict__ qkv,        // (B, H, {q,k,v}, DH): HF's packed projection
                      const T* __restrict__ cs, const T* __restrict__ sn,   // (B, rot)
                      T* __restrict__ kc, T* __restrict__ vc,               // (B, H, Tmax, DH)
                      int32_t* __restrict__ indir,                          // (B, Tmax): cache row of position t
                      const T* __restrict__ add_mask,                       // (B, Tmax): 0 or -inf
                      const int64_t* __restrict__ cursor,                   // slot of the new token
                      T* __restrict__ out,                                  // (B, H*DH)
                      int H, int Tmax, int rot, float scale) {
  constexpr int N = Vec16<T>::N, VEC = DH / N, G = LD_THREADS / VEC, CH = 10;
  extern __shared__ float dsm[];            // probabilities [Tmax] | cache rows [Tmax]
  float* ssc = dsm;
  int* srow = reinterpret_cast<int*>(dsm + Tmax);
  __shared__ float sq[DH], sk[DH], sv[DH], sred[32];
  __shared__ float sacc[G * DH];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();                               // qkv, the cursor, indir and the caches all come from earlier kernels
  const int cur = (int)*cursor;
  if (cur < 0 || cur >= Tmax) return;       // the host sizes the caches for prompt + max_new_tokens
  const int half = rot >> 1;
  const T* src = qkv + ((int64_t)b * H + h) * 3 * DH;
  if (tid < DH) {
    const int d = tid;
    float q = Elem<T>::to_f(src[d]), k = Elem<T>::to_f(src[DH + d]);
    const T v = src[2 * DH + d];
    if (d < rot) {
      // q*cos + rotate_half(q)*sin: first half x1 c1 - x2 s1, second half x2 c2 + x1 s2
      const int dp = d < half ? d + half : d - half;
      const float c = Elem<T>::to_f(cs[(int64_t)b * rot + d]), s = Elem<T>::to_f(sn[(int64_t)b * rot + d]);
      const float qp = Elem<T>::to_f(src[dp]), kp = Elem<T>::to_f(src[DH + dp]);
      q = d < half ? q * c - qp * s : q * c + qp * s;
      k = d < half ? k * c - kp * s : k * c + kp * s;
    }
    const T qr = Elem<T>::from_f(q), kr = Elem<T>::from_f(k);   // rounded to the storage type, as prefill does
    sq[d] = Elem<T>::to_f(qr);
    sk[d] = Elem<T>::to_f(kr);
    sv[d] = Elem<T>::to_f(v);
    const int64_t slot = (((int64_t)b * H + h) * Tmax + cur) * DH + d;
    kc[slot] = kr;
    vc[slot] = v;
    if (h == 0 && d == 0) indir[(int64_t)b * Tmax + cur] = b;
  }
  __syncthreads();
  // ---- scores: one cached key per thread (its whole row: VEC independent 16-byte loads), two keys'
  // worth of loads in flight per thread and pass
  float m = -INFINITY;
  for (int j0 = tid; j0 < cur; j0 += 2 * LD_THREADS) {
    const int j1 = j0 + LD_THREADS;
    const bool two = j1 < cur;
    const int row0 = indir[(int64_t)b * Tmax + j0], row1 = two ? indir[(int64_t)b * Tmax + j1] : row0;
    const T* kp0 = kc + (((int64_t)row0 * H + h) * Tmax + j0) * DH;
    const T* kp1 = kc + (((int64_t)row1 * H + h) * Tmax + (two ? j1 : j0)) * DH;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c0 = 0; c0 < VEC; c0 += CH) {
      Vec16<T> ka[CH], kb[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (c0 + c < VEC) {
          ka[c].load(kp0 + (c0 + c) * N);
          kb[c].load(kp1 + (c0 + c) * N);
        }
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (c0 + c < VEC) {
          float f[N], e[N];
          ka[c].unpack(f);
          kb[c].unpack(e);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            s0 = fmaf(f[i], sq[(c0 + c) * N + i], s0);
            s1 = fmaf(e[i], sq[(c0 + c) * N + i], s1);
          }
        }
    }
    s0 = s0 * scale + Elem<T>::to_f(add_mask[(int64_t)b * Tmax + j0]);
    ssc[j0] = s0;
    srow[j0] = row0;
    m = fmaxf(m, s0);
    if (two) {
      s1 = s1 * scale + Elem<T>::to_f(add_mask[(int64_t)b * Tmax + j1]);
      ssc[j1] = s1;
      srow[j1] = row1;
      m = fmaxf(m, s1);
    }
  }
  if (tid == 0) {                           // the new token's own key never leaves the SM
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) s = fmaf(sq[d], sk[d], s);
    s = s * scale + Elem<T>::to_f(add_mask[(int64_t)b * Tmax + cur]);
    ssc[cur] = s;
    m = fmaxf(m, s);
  }
  m = block_max(m, sred);
  float l = 0.f;
  for (int j = tid; j <= cur; j += LD_THREADS) {
    const float p = __expf(ssc[j] - m);
    ssc[j] = p;
    l += p;
  }
  l = block_sum(l, sred);                   // (its barriers also publish ssc)
  // ---- PV: VEC threads per cached row (coalesced 16-byte loads), G rows in flight per pass ----------
  const int g = tid / VEC, c = tid - g * VEC;
  if (g < G) {
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    constexpr int PB = 8;                   // cached rows in flight per thread and pass
    for (int j0 = g; j0 < cur; j0 += PB * G) {
      Vec16<T> vv[PB];
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        const int j = j0 + u * G;
        if (j < cur) vv[u].load(vc + (((int64_t)srow[j] * H + h) * Tmax + j) * DH + c * N);
      }
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        const int j = j0 + u * G;
        if (j < cur) {
          float f[N];
          vv[u].unpack(f);
          const float p = ssc[j];
#pragma unroll
          for (int i = 0; i < N; ++i) acc[i] = fmaf(p, f[i], acc[i]);
        }
      }
    }
    if (g == 0) {
      const float p = ssc[cur];
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = fmaf(p, sv[c * N + i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) sacc[g * DH + c * N + i] = acc[i];
  }
  __syncthreads();
  if (tid < DH) {
    float s = 0.f;
    for (int gg = 0; gg < G; ++gg) s += sacc[gg * DH + tid];   // fixed order
    out[((int64_t)b * H + h) * DH + tid] = Elem<T>::from_f(s / l);
  }
}

template <typename T, int DH>
static int launch_lm_decode_attn(const void* qkv, const void* cs, const void* sn, void* kc, void* vc,
                                 int32_t* indir, const void* add_mask, const int64_t* cursor, void* out,
                                 int B, int H, int Tmax, int rot, float scale, cudaStream_t st) {
  const size_t smem = (size_t)Tmax * 8;
  cudaError_t e = launch_pdl(lm_decode_attn_kernel<T, DH>, dim3(H, B), dim3(LD_THREADS), smem, st, (const T*)qkv,
                             (const T*)cs, (const T*)sn, (T*)kc, (T*)vc, indir, (const T*)add_mask, cursor,
                             (T*)out, H, Tmax, rot, scale);
  if (e != cudaSuccess) { set_error("lm_decode_attn launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// skinny linear
// ------------------------------------------------------------------------------------------------
constexpr int SM_ROWS = 16;

__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream16(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// CTA = RT tiles of 16 weight rows x all of K; warp w takes the 32-wide K steps w, w+WARPS, ...  Lane
// (g = lane/4, t = lane%4) loads 16 bytes (8 consecutive k) of weight rows g and g+8 of each tile and
// of x row g: the eight values fill the k slots {2t,2t+1,2t+8,2t+9} of TWO m16n8k16 MMAs.  A dot
// product does not care in which order its terms are added, so the A (weights) and B (x) fragments
// only have to agree on the slot -> k assignment, and both come from the same 16-byte chunk index.
// The weights do not depend on the previous kernel: the first batch of weight loads is issued BEFORE
// the programmatic-dependency wait (it streams while the producer of x drains).
// U = K steps per batch (loads in flight per lane: U x RT x 32 bytes).
// KS > 1: split-K over a thread-block cluster of KS CTAs (grid.y): each CTA streams 1/KS of K for the
// same row tile, the partial sums meet in CTA 0's shared memory (distributed shared memory stores, one
// cluster barrier) and CTA 0 runs the epilogue — for the projections with few rows and a long K
// (4h -> h: 160 row tiles x 20 KB rows), which otherwise leave most SMs with one CTA's loads in flight.
template <int ACT, int WARPS, int U, int RT, int KS>
__global__ void __launch_bounds__(WARPS * 32)
linear_small_m_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                      const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y, int M, int N,
                      int K) {
  __shared__ float red[WARPS][RT * SM_ROWS][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * (RT * SM_ROWS);
  // rows past N / beams past M read a valid row instead and are dropped at the store
  const uint4 *wa[RT], *wb[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    wa[r] = reinterpret_cast<const uint4*>(W + (int64_t)min(n0 + r * SM_ROWS + g, N - 1) * K) + t;
    wb[r] = reinterpret_cast<const uint4*>(W + (int64_t)min(n0 + r * SM_ROWS + g + 8, N - 1) * K) + t;
  }
  const uint4* xp = reinterpret_cast<const uint4*>(x + (int64_t)min(g, M - 1) * K) + t;
  int steps = K >> 5;
  unsigned ks = 0;
  if (KS > 1) {                             // this CTA's share of the K steps
    // a CTA may only write into a peer's shared memory once that peer has started: everybody arrives at
    // the cluster barrier here (non-blocking) and waits for it right before the remote stores
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(ks));
    const int lo = (int)(((int64_t)steps * ks) / KS), hi = (int)(((int64_t)steps * (ks + 1)) / KS);
#pragma unroll
    for (int r = 0; r < RT; ++r) { wa[r] += lo * 4; wb[r] += lo * 4; }
    xp += lo * 4;
    steps = hi - lo;
  }
  constexpr int STRIDE = WARPS * U;
  float acc[RT][4];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
  pdl_launch_dependents();
  {                                         // first batch: weights before the dependency wait, x after it
    uint4 a[RT][U], bq[RT][U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = warp + u * WARPS;
      if (s < steps) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          a[r][u] = ldg_stream16(wa[r] + s * 4);
          bq[r][u] = ldg_stream16(wb[r] + s * 4);
        }
      }
    }
    pdl_wait();                             // x (and the buffer y may reuse) belong to earlier kernels
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = warp + u * WARPS;
      if (s < steps) xv[u] = __ldg(xp + s * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = warp + u * WARPS;
      if (s < steps) {                      // warp-uniform
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          mma_16816(acc[r], a[r][u].x, bq[r][u].x, a[r][u].y, bq[r][u].y, xv[u].x, xv[u].y);
          mma_16816(acc[r], a[r][u].z, bq[r][u].z, a[r][u].w, bq[r][u].w, xv[u].z, xv[u].w);
        }
      }
    }
  }
  for (int s0 = warp + STRIDE; s0 < steps; s0 += STRIDE) {
    uint4 a[RT][U], bq[RT][U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u * WARPS;
      if (s < steps) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          a[r][u] = ldg_stream16(wa[r] + s * 4);
          bq[r][u] = ldg_stream16(wb[r] + s * 4);
        }
        xv[u] = __ldg(xp + s * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u * WARPS;
      if (s < steps) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          mma_16816(acc[r], a[r][u].x, bq[r][u].x, a[r][u].y, bq[r][u].y, xv[u].x, xv[u].y);
          mma_16816(acc[r], a[r][u].z, bq[r][u].z, a[r][u].w, bq[r][u].w, xv[u].z, xv[u].w);
        }
      }
    }
  }
  // D fragment: c0 = (row g, beam 2t), c1 = (g, 2t+1), c2 = (g+8, 2t), c3 = (g+8, 2t+1)
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    red[warp][r * SM_ROWS + g][2 * t] = acc[r][0];
    red[warp][r * SM_ROWS + g][2 * t + 1] = acc[r][1];
    red[warp][r * SM_ROWS + g + 8][2 * t] = acc[r][2];
    red[warp][r * SM_ROWS + g + 8][2 * t + 1] = acc[r][3];
  }
  __syncthreads();
  __shared__ float xchg[KS > 1 ? KS : 1][RT * SM_ROWS * 8];
  if (KS > 1) {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");    // every CTA of the cluster is running
    for (int o = tid; o < RT * SM_ROWS * 8; o += WARPS * 32) {
      const int r = o % (RT * SM_ROWS), mrow = o / (RT * SM_ROWS);
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) s += red[w][r][mrow];   // fixed order
      uint32_t local = (uint32_t)__cvta_generic_to_shared(&xchg[ks][o]), remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (ks != 0) return;
  }
  for (int o = tid; o < RT * SM_ROWS * 8; o += WARPS * 32) {
    const int r = o % (RT * SM_ROWS), mrow = o / (RT * SM_ROWS);
    const int n = n0 + r;
    if (mrow < M && n < N) {
      float s = 0.f;
      if (KS > 1) {
#pragma unroll
        for (int c = 0; c < KS; ++c) s += xchg[c][o];         // fixed order
      } else {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += red[w][r][mrow];   // fixed order
      }
      if (bias) s += __bfloat162float(bias[n]);
      if (ACT == 1) s = 0.5f * s * (1.f + erff(s * 0.70710678118654752f));   // exact GELU
      y[(int64_t)mrow * N + n] = __float2bfloat16_rn(s);
    }
  }
}

template <int WARPS, int U, int RT, int KS>
static cudaError_t launch_linear_small_m(int act, cudaStream_t st, const __nv_bfloat16* x, const __nv_bfloat16* w,
                                         const __nv_bfloat16* bias, __nv_bfloat16* y, int M, int N, int K) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((N + RT * SM_ROWS - 1) / (RT * SM_ROWS)), KS);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (KS > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 1;
    at[na].val.clusterDim.y = KS;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return act == 1 ? cudaLaunchKernelEx(&cfg, linear_small_m_kernel<1, WARPS, U, RT, KS>, x, w, bias, y, M, N, K)
                  : cudaLaunchKernelEx(&cfg, linear_small_m_kernel<0, WARPS, U, RT, KS>, x, w, bias, y, M, N, K);
}

// ------------------------------------------------------------------------------------------------
// beam search: log-softmax + running score + top-K over (beams x vocabulary) per batch item
// ------------------------------------------------------------------------------------------------
// HF `_beam_search` per step: `log_softmax(logits)`, `+ running_beam_scores[:, :, None]`, `topk(2 * num_beams)`
// over the flattened (beams * V) axis — in torch a softmax kernel, an add and an 8-kernel radix top-k over
// 370 k floats (~130 us of a 2.6 ms token).  Here: one pass over the logits (each CTA: a 4096-wide chunk of
// one row -> chunk max, chunk sum of exponentials, chunk top-K by K rounds of block arg-max in shared
// memory), then one small merge per batch item.  Ties break towards the smaller flat index.
constexpr int BT_THREADS = 256, BT_CHUNK = 4096, BT_KMAX = 16;

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax argmax_pick(ArgMax a, ArgMax b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, ArgMax* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = argmax_pick(a, b);
  }
  __syncthreads();
  if (lane == 0) sh[w] = a;
  __syncthreads();
  ArgMax r = sh[0];
#pragma unroll
  for (int k = 1; k < BT_THREADS / 32; ++k) r = argmax_pick(r, sh[k]);
  return r;
}

// grid (splits, rows); partial layout per (row, split): [max, sum, K values, K token ids (as float bits)]
__global__ void __launch_bounds__(BT_THREADS)
beam_partial_kernel(const float* __restrict__ logits, int64_t ld, int V, int K, float* __restrict__ part) {
  __shared__ float sx[BT_CHUNK];
  __shared__ float sred[32];
  __shared__ ArgMax sam[BT_THREADS / 32];
  const int row = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const int c0 = split * BT_CHUNK, n = min(BT_CHUNK, V - c0);
  const float* x = logits + (int64_t)row * ld + c0;
  float m = -INFINITY;
  for (int i = tid; i < n; i += BT_THREADS) {
    const float v = x[i];
    sx[i] = v;
    m = fmaxf(m, v);
  }
  m = block_max(m, sred);
  float s = 0.f;
  if (m > -INFINITY)
    for (int i = tid; i < n; i += BT_THREADS) s += __expf(sx[i] - m);
  s = block_sum(s, sred);
  float* out = part + ((int64_t)row * gridDim.x + split) * (2 + 2 * BT_KMAX);
  if (tid == 0) { out[0] = m; out[1] = s; }
  for (int k = 0; k < K; ++k) {
    ArgMax a = {-INFINITY, 0x7fffffff};
    for (int i = tid; i < n; i += BT_THREADS) {
      ArgMax b = {sx[i], i};
      a = argmax_pick(a, b);
    }
    a = block_argmax(a, sam);
    if (tid == 0) {
      out[2 + k] = a.v;
      out[2 + BT_KMAX + k] = __int_as_float(a.i < n ? c0 + a.i : -1);
      if (a.i < n) sx[a.i] = -INFINITY;        // taken
    }
    __syncthreads();
  }
}

// grid (B); one batch item: nb rows x splits partials -> top-K of log_softmax + running score
__global__ void __launch_bounds__(BT_THREADS)
beam_merge_kernel(const float* __restrict__ part, const float* __restrict__ running, int nb, int splits, int V,
                  int K, float* __restrict__ top_lp, int64_t* __restrict__ top_idx) {
  extern __shared__ float dsm2[];            // candidate scores [nb*splits*K] | flat ids (int) [nb*splits*K]
  __shared__ float lse[64];
  __shared__ ArgMax sam[BT_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ncand = nb * splits * K;
  float* cv = dsm2;
  int* ci = reinterpret_cast<int*>(dsm2 + ncand);
  const int stride = 2 + 2 * BT_KMAX;
  if (tid < nb) {                            // log-sum-exp of the row from its chunk partials (fixed order)
    const float* p = part + (int64_t)(b * nb + tid) * splits * stride;
    float M = -INFINITY;
    for (int s = 0; s < splits; ++s) M = fmaxf(M, p[s * stride]);
    float S = 0.f;
    for (int s = 0; s < splits; ++s) S += p[s * stride + 1] * __expf(p[s * stride] - M);
    lse[tid] = M + logf(S);
  }
  __syncthreads();
  for (int c = tid; c < ncand; c += BT_THREADS) {
    const int r = c / (splits * K), rem = c - r * splits * K, s = rem / K, k = rem - s * K;
    const float* p = part + ((int64_t)(b * nb + r) * splits + s) * stride;
    const int tok = __float_as_int(p[2 + BT_KMAX + k]);
    cv[c] = tok >= 0 ? (p[2 + k] - lse[r]) + running[b * nb + r] : -INFINITY;
    ci[c] = tok >= 0 ? r * V + tok : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    ArgMax a = {-INFINITY, 0x7fffffff};
    int pos = -1;
    for (int c = tid; c < ncand; c += BT_THREADS) {
      ArgMax bb = {cv[c], ci[c]};
      const ArgMax n2 = argmax_pick(a, bb);
      if (n2.i != a.i || n2.v != a.v) pos = c;
      a = n2;
    }
    const ArgMax w = block_argmax(a, sam);
    if (pos >= 0 && a.i == w.i && a.v == w.v) cv[pos] = -INFINITY;     // the winner's owner retires it (ids are unique)
    if (tid == 0) {
      top_lp[b * K + k] = w.v;
      top_idx[b * K + k] = w.i;
    }
    __syncthreads();
  }
}

}  // namespace unimp

using namespace unimp;

extern "C" int64_t unimp_beam_topk_workspace(int rows, int V) {
  const int splits = (V + BT_CHUNK - 1) / BT_CHUNK;
  return (int64_t)rows * splits * (2 + 2 * BT_KMAX) * sizeof(float);
}

extern "C" int unimp_beam_topk(const float* logits, int64_t ld, const float* running, int B, int nb, int V, int K,
                               void* workspace, float* top_lp, int64_t* top_idx, void* stream) {
  UNIMP_CHECK_ARG(logits && running && workspace && top_lp && top_idx, UNIMP_E_NULL, "beam_topk: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && nb > 0 && nb <= 64 && V > 0 && K > 0 && K <= BT_KMAX && (int64_t)nb * V < 0x7fffffff &&
                      K <= V,
                  UNIMP_E_SHAPE, "beam_topk: need 1 <= K <= %d, K <= V, beams <= 64, beams * V < 2^31", BT_KMAX);
  const int splits = (V + BT_CHUNK - 1) / BT_CHUNK;
  const size_t smem = (size_t)nb * splits * K * 8;
  UNIMP_CHECK_ARG(smem <= 48 * 1024, UNIMP_E_SHAPE, "beam_topk: beams * ceil(V / 4096) * K = %d candidates exceed 6144",
                  nb * splits * K);
  cudaStream_t st = (cudaStream_t)stream;
  beam_partial_kernel<<<dim3(splits, B * nb), BT_THREADS, 0, st>>>(logits, ld, V, K, (float*)workspace);
  UNIMP_CHECK_LAUNCH();
  beam_merge_kernel<<<B, BT_THREADS, smem, st>>>((const float*)workspace, running, nb, splits, V, K, top_lp, top_idx);
  UNIMP_CHECK_LAUNCH();
  return 0;
}


extern "C" int unimp_lm_decode_attn(const void* qkv, const void* cos, const void* sin, void* k_cache,
                                    void* v_cache, int32_t* indir, const void* add_mask,
                                    const int64_t* cursor, void* out, int B, int H, int Tmax, int dh, int rot,
                                    float scale, int dtype, void* stream) {
  UNIMP_CHECK_ARG(qkv && cos && sin && k_cache && v_cache && indir && add_mask && cursor && out, UNIMP_E_NULL,
                  "lm_decode_attn: NULL pointer");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "lm_decode_attn: dtype");
  UNIMP_CHECK_ARG(B > 0 && H > 0 && Tmax > 0 && Tmax <= 5120, UNIMP_E_SHAPE,
                  "lm_decode_attn: bad B/H/Tmax (Tmax=%d must be in 1..5120)", Tmax);
  UNIMP_CHECK_ARG(rot >= 0 && rot <= dh && rot % 2 == 0, UNIMP_E_SHAPE,
                  "lm_decode_attn: rotary_dim=%d must be even and <= head_dim=%d", rot, dh);
  UNIMP_CHECK_ARG(aligned16(k_cache) && aligned16(v_cache), UNIMP_E_ALIGN,
                  "lm_decode_attn: the caches must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
#define UNIMP_LD_CASE(DHV)                                                                                   \
  case DHV:                                                                                                  \
    return dtype == UNIMP_BF16                                                                               \
               ? launch_lm_decode_attn<__nv_bfloat16, DHV>(qkv, cos, sin, k_cache, v_cache, indir, add_mask, \
                                                           cursor, out, B, H, Tmax, rot, scale, st)          \
               : launch_lm_decode_attn<float, DHV>(qkv, cos, sin, k_cache, v_cache, indir, add_mask, cursor, \
                                                   out, B, H, Tmax, rot, scale, st)
  switch (dh) {
    UNIMP_LD_CASE(32);
    UNIMP_LD_CASE(64);
    UNIMP_LD_CASE(80);
    UNIMP_LD_CASE(96);
    UNIMP_LD_CASE(128);
    default:
      set_error("lm_decode_attn: head_dim=%d not built (32, 64, 80, 96, 128)", dh);
      return UNIMP_E_SHAPE;
  }
#undef UNIMP_LD_CASE
}

extern "C" int unimp_linear_small_m(const void* x, const void* w, const void* bias, void* y, int M, int N,
                                    int K, int act, int dtype, void* stream) {
  UNIMP_CHECK_ARG(x && w && y, UNIMP_E_NULL, "linear_small_m: NULL pointer");
  UNIMP_CHECK_ARG(dtype == UNIMP_BF16, UNIMP_E_DTYPE, "linear_small_m: bf16 only");
  UNIMP_CHECK_ARG(M >= 1 && M <= 8, UNIMP_E_SHAPE, "linear_small_m: M=%d rows (1..8: the beams of one decode step)", M);
  UNIMP_CHECK_ARG(N >= 1 && K >= 32 && K % 32 == 0, UNIMP_E_SHAPE,
                  "linear_small_m: K=%d must be a positive multiple of 32", K);
  UNIMP_CHECK_ARG(act == 0 || act == 1, UNIMP_E_SHAPE, "linear_small_m: act must be 0 (none) or 1 (exact GELU)");
  UNIMP_CHECK_ARG(aligned16(x) && aligned16(w), UNIMP_E_ALIGN, "linear_small_m: x / w must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16 *xb = (const __nv_bfloat16*)x, *wb = (const __nv_bfloat16*)w, *bb = (const __nv_bfloat16*)bias;
  __nv_bfloat16* yb = (__nv_bfloat16*)y;
  static int cfg = -1;                             // UNIMP_GEMV_CFG=[<ks>]<warps><u><rt> (e.g. 841, 4441): experiments
  if (cfg < 0) { const char* ev = getenv("UNIMP_GEMV_CFG"); cfg = ev ? atoi(ev) : 0; }
  int c = cfg;
  // measured per shape on a B200 (tools/gemv_bench.py, cold weights; profiles/r2_decode_gemv_sweep.log):
  // with >= 2 row tiles per SM small batches at high occupancy win (48 registers, 8 CTAs / SM); few
  // tiles over a long K (4h -> h) want the K range split over a 4-CTA cluster; a handful of tiles
  // (to_q: N = 512) likewise; the rest (N = 2560, K <= 2560) deep batches in one CTA per tile
  if (!c) {
    const int tiles = (N + SM_ROWS - 1) / SM_ROWS;
    c = tiles >= 2 * UNIMP_NUM_SMS ? 821 : (N <= 1024 ? 4841 : (K >= 4096 ? 4441 : 881));
  }
  cudaError_t e;
  switch (c) {
#define UNIMP_SM_CASE(code, WARPS, U, RT, KS) \
    case code: e = launch_linear_small_m<WARPS, U, RT, KS>(act, st, xb, wb, bb, yb, M, N, K); break
    UNIMP_SM_CASE(821, 8, 2, 1, 1);
    UNIMP_SM_CASE(881, 8, 8, 1, 1);
    UNIMP_SM_CASE(841, 8, 4, 1, 1);
    UNIMP_SM_CASE(1641, 16, 4, 1, 1);
    UNIMP_SM_CASE(4841, 8, 4, 1, 4);      // 4-CTA cluster split-K
    UNIMP_SM_CASE(4441, 4, 4, 1, 4);
#undef UNIMP_SM_CASE
    default:
      set_error("linear_small_m: unknown UNIMP_GEMV_CFG=%d", c);
      return UNIMP_E_SHAPE;
  }
  if (e != cudaSuccess) { set_error("linear_small_m launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
