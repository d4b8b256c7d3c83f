// bf16 attention cores on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, operands
// staged by TMA (SWIZZLE_128B), softmax on the 128 threads that own the 128 TMEM lanes.
//
//   K1  masked media-located cross-attention (MaskedCrossAttention core):
//       a 128-query tile walks only the 64-key image blocks its rows reference
//       (text_time-1); the mask is evaluated in registers, never materialised.
//   K2  Perceiver latent attention (64 x 320) and K3 ViT-L/14 self-attention (257 x 257):
//       all of S = Q K^T for the tile lives in TMEM (<= 384 fp32 columns), so softmax is a
//       plain two-pass row softmax with no online rescaling of O.
//
// One CTA = one (128-row query tile, head, batch).  Thread t owns query row t = TMEM lane t.
// Thread 0 additionally issues every TMA and MMA (tcgen05.mma is a single-thread instruction).
//
// Roofline: fwd FLOPs = 4*dh*Lq*Lk_attended per (b,h); bytes = Q + O + K + V (+lse).
#include "common.cuh"
#include "tc_common.cuh"

namespace unimp {

using namespace tc;

constexpr int TQ = 128;        // query rows per CTA (UMMA M)
constexpr int KB = 64;         // keys per block (UMMA N for S, K-extent for PV)
constexpr int DH = 64;
constexpr int MAX_BLOCKS_UNMASKED = 6;       // 384 fp32 S columns + 64 O columns <= 512
constexpr uint32_t Q_BYTES = TQ * DH * 2;    // 16 KB
constexpr uint32_t KV_BYTES = KB * DH * 2;   // 8 KB
constexpr uint32_t P_BYTES = TQ * KB * 2;    // 16 KB

struct FwdArgs {
  __nv_bfloat16* o;
  int64_t o_bs, o_rs;
  float* lse;
  const int32_t* tt;
  int Lq, Lk, H, n, Ti;
  float scale, scale_log2;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// Writes this thread's 64 probabilities (two halves of 32) as one 128-byte swizzled row of sP.
__device__ __forceinline__ void store_p_half(uint8_t* sP, int row, int half, const float* p) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    v.x = pack_bf16(p[8 * c + 0], p[8 * c + 1]);
    v.y = pack_bf16(p[8 * c + 2], p[8 * c + 3]);
    v.z = pack_bf16(p[8 * c + 4], p[8 * c + 5]);
    v.w = pack_bf16(p[8 * c + 6], p[8 * c + 7]);
    *reinterpret_cast<uint4*>(sP + sw128_offset(row, half * 4 + c)) = v;
  }
}

constexpr int FWD_THREADS = TQ + 32;  // warps 0-3: one query row per thread; warp 4: TMA/MMA issuer

template <bool MASKED>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const FwdArgs a) {
  constexpr int NSLOT = MASKED ? 1 : MAX_BLOCKS_UNMASKED;
  constexpr uint32_t S_COL = 0;
  constexpr uint32_t O_COL = MASKED ? KB : MAX_BLOCKS_UNMASKED * KB;
  constexpr uint32_t TMEM_COLS = MASKED ? 128 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + NSLOT * KV_BYTES;
  uint8_t* sP = sV + NSLOT * KV_BYTES;  // two P buffers
  __shared__ uint64_t bar_q, bar_k, bar_v, bar_s, bar_p[2], bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_jlo, s_jhi;

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool issuer = tid == TQ;        // lane 0 of warp 4
  const bool worker = tid < TQ;         // owns TMEM lane `tid`
  const int row0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int row = row0 + tid;
  const bool valid = worker && row < a.Lq;

  if (issuer) {
    mbar_init(&bar_q, 1); mbar_init(&bar_k, 1); mbar_init(&bar_v, 1); mbar_init(&bar_s, 1);
    mbar_init(&bar_p[0], 1); mbar_init(&bar_p[1], 1); mbar_init(&bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv);
    s_jlo = 1 << 30; s_jhi = -1;
  }
  if (warp == 4) tmem_alloc(&tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  if (issuer) {
    mbar_arrive_expect_tx(&bar_q, Q_BYTES);
    tma_load_4d(sQ, &tq, &bar_q, 0, h, row0, b);
  }

  const uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);   // S = Q K^T : A, B K-major
  const uint32_t idesc_o = make_idesc(TQ, DH, 0, 1);   // O = P V   : A K-major, B (V) MN-major
  float sum = 0.f, lse_val = -INFINITY;
  uint32_t r[32];
  bool any_mma = true;

  if constexpr (!MASKED) {
    const int nb = (a.Lk + KB - 1) / KB;
    if (issuer) {
      mbar_arrive_expect_tx(&bar_k, nb * KV_BYTES);
      for (int j = 0; j < nb; ++j) tma_load_4d(sK + j * KV_BYTES, &tk, &bar_k, 0, h, j * KB, b);
      mbar_arrive_expect_tx(&bar_v, nb * KV_BYTES);
      for (int j = 0; j < nb; ++j) tma_load_4d(sV + j * KV_BYTES, &tv, &bar_v, 0, h, j * KB, b);
      mbar_wait(&bar_q, 0);
      mbar_wait(&bar_k, 0);
      tcgen05_fence_after();
      for (int j = 0; j < nb; ++j) {
#pragma unroll
        for (int k4 = 0; k4 < DH / 16; ++k4)
          umma_ss(tmem + S_COL + j * KB, make_smem_desc(smem_u32(sQ) + k4 * 32, 16, 1024),
                  make_smem_desc(smem_u32(sK + j * KV_BYTES) + k4 * 32, 16, 1024), idesc_s, k4 > 0);
      }
      umma_commit(&bar_s);
    }
    float m = -INFINITY, ms = 0.f;
    if (worker) {
      mbar_wait(&bar_s, 0);
      tcgen05_fence_after();
      // pass 1: row max over all keys
      for (int j = 0; j < nb; ++j) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(lane_addr + S_COL + j * KB + half * 32, r);
          tmem_ld_wait();
          const int c0 = j * KB + half * 32;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c0 + c < a.Lk) m = fmaxf(m, __uint_as_float(r[c]));
        }
      }
      ms = m * a.scale_log2;
    }
    // pass 2: exponentiate, write P_j, issue O += P_j V_j
    for (int j = 0; j < nb; ++j) {
      const int pb = j & 1;
      if (worker) {
        if (j >= 2) mbar_wait(&bar_p[pb], ((j >> 1) - 1) & 1);  // PV of block j-2 released sP[pb]
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(lane_addr + S_COL + j * KB + half * 32, r);
          tmem_ld_wait();
          const int c0 = j * KB + half * 32;
          float p[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            p[c] = (c0 + c < a.Lk) ? exp2f(__uint_as_float(r[c]) * a.scale_log2 - ms) : 0.f;
            sum += p[c];
          }
          store_p_half(sP + pb * P_BYTES, tid, half, p);
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
      }
      __syncthreads();
      if (issuer) {
        if (j == 0) mbar_wait(&bar_v, 0);
        tcgen05_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < KB / 16; ++k4)
          umma_ss(tmem + O_COL, make_smem_desc(smem_u32(sP + pb * P_BYTES) + k4 * 32, 16, 1024),
                  make_smem_desc(smem_u32(sV + j * KV_BYTES) + k4 * 2048, 1024, 1024), idesc_o,
                  (j > 0 || k4 > 0));
        umma_commit(&bar_p[pb]);
        if (j == nb - 1) umma_commit(&bar_o);
      }
    }
    lse_val = m * a.scale + logf(sum);
  } else {
    // ---- masked: which image blocks does this tile touch? ------------------------------
    int ttr = 0;
    if (valid) ttr = a.tt[(int64_t)b * a.Lq + row];
    const bool uniform = ttr > a.Ti;
    const int blk = (ttr >= 1 && !uniform) ? ttr - 1 : -1;
    if (uniform) { atomicMin(&s_jlo, 0); atomicMax(&s_jhi, a.Ti - 1); }
    else if (blk >= 0) { atomicMin(&s_jlo, blk); atomicMax(&s_jhi, blk); }
    __syncthreads();
    const int jlo = s_jlo, jhi = s_jhi;
    any_mma = jhi >= jlo;
    if (!any_mma && issuer) mbar_wait(&bar_q, 0);  // never leave a TMA in flight at exit
    float m_row = 0.f;
    for (int j = jlo; j <= jhi; ++j) {
      const int it = j - jlo;
      const uint32_t ph = it & 1;
      if (issuer) {
        // the previous PV MMA (reads sK/sV/sP) has completed: waited on bar_p below
        mbar_arrive_expect_tx(&bar_k, 2 * KV_BYTES);
        tma_load_4d(sK, &tk, &bar_k, 0, h, j * a.n, b);
        tma_load_4d(sV, &tv, &bar_k, 0, h, j * a.n, b);
        if (it == 0) mbar_wait(&bar_q, 0);
        mbar_wait(&bar_k, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < DH / 16; ++k4)
          umma_ss(tmem + S_COL, make_smem_desc(smem_u32(sQ) + k4 * 32, 16, 1024),
                  make_smem_desc(smem_u32(sK) + k4 * 32, 16, 1024), idesc_s, k4 > 0);
        umma_commit(&bar_s);
      }
      if (worker) {
        mbar_wait(&bar_s, ph);
        tcgen05_fence_after();
        const bool mine = uniform || blk == j;
        float sv[64];
        tmem_ld32(lane_addr + S_COL, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) sv[c] = __uint_as_float(r[c]);
        tmem_ld32(lane_addr + S_COL + 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) sv[32 + c] = __uint_as_float(r[c]);
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; ++c) m = fmaxf(m, sv[c]);
        if (uniform) m = 0.f;
        const float ms = m * a.scale_log2;
        float psum = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float p = uniform ? 1.f : exp2f(sv[c] * a.scale_log2 - ms);
          p = mine ? p : 0.f;
          sv[c] = p;
          psum += p;
        }
        if (mine) { sum += psum; m_row = m; }
        store_p_half(sP, tid, 0, sv);
        store_p_half(sP, tid, 1, sv + 32);
        fence_proxy_async_smem();
        tcgen05_fence_before();
      }
      __syncthreads();
      if (issuer) {
        tcgen05_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < KB / 16; ++k4)
          umma_ss(tmem + O_COL, make_smem_desc(smem_u32(sP) + k4 * 32, 16, 1024),
                  make_smem_desc(smem_u32(sV) + k4 * 2048, 1024, 1024), idesc_o, (it > 0 || k4 > 0));
        umma_commit(&bar_p[0]);
        if (j == jhi) umma_commit(&bar_o);
      }
      mbar_wait(&bar_p[0], ph);  // sK/sV/sP free again; S may be overwritten
    }
    lse_val = sum > 0.f ? m_row * a.scale + logf(sum) : -INFINITY;
  }

  // ---- epilogue: O / sum -> global -----------------------------------------------------
  if (worker) {
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (any_mma) {
      mbar_wait(&bar_o, 0);
      tcgen05_fence_after();
    }
    __nv_bfloat16* orow = a.o + (int64_t)b * a.o_bs + (int64_t)row * a.o_rs + h * DH;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (any_mma) {
        tmem_ld32(lane_addr + O_COL + half * 32, r);
        tmem_ld_wait();
      }
      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (any_mma && inv > 0.f) {
            v.x = pack_bf16(__uint_as_float(r[8 * c + 0]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
          }
          *reinterpret_cast<uint4*>(orow + half * 32 + c * 8) = v;
        }
      }
    }
    if (valid) a.lse[((int64_t)b * a.H + h) * a.Lq + row] = lse_val;
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
}

// ---- host ---------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bhld(CUtensorMap* out, const void* base, int64_t batch_stride, int64_t row_stride, int B,
                   int L, int H, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return UNIMP_E_DEVICE; }
  if (B == 1) batch_stride = row_stride * (int64_t)L;  // a size-1 dim may carry any stride
  cuuint64_t dims[4] = {(cuuint64_t)DH, (cuuint64_t)H, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)DH * 2, (cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)DH, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p bs=%lld rs=%lld B=%d L=%d H=%d", (int)r, base,
              (long long)batch_stride, (long long)row_stride, B, L, H);
    return UNIMP_E_SHAPE;
  }
  return 0;
}

static bool view_ok(const void* p, int64_t bs, int64_t rs) {
  return aligned16(p) && (bs % 8 == 0) && (rs % 8 == 0);
}

bool attn_fwd_tc_supported(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_mview_t o,
                           const int32_t* tt, int Lq, int Lk, int n, int dh) {
  (void)Lq;
  if (dh != DH) return false;
  if (!view_ok(q.ptr, q.batch_stride, q.row_stride) || !view_ok(k.ptr, k.batch_stride, k.row_stride) ||
      !view_ok(v.ptr, v.batch_stride, v.row_stride) || !view_ok(o.ptr, o.batch_stride, o.row_stride))
    return false;
  if (tt) return n == KB;
  return Lk <= MAX_BLOCKS_UNMASKED * KB;
}

template <bool MASKED>
static int launch_fwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt, unimp_mview_t o,
                      float* lse, int B, int Lq, int Lk, int H, int n, int Ti, float scale,
                      cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bhld(&tq, q.ptr, q.batch_stride, q.row_stride, B, Lq, H, TQ))) return rc;
  if ((rc = make_tmap_bhld(&tk, k.ptr, k.batch_stride, k.row_stride, B, Lk, H, KB))) return rc;
  if ((rc = make_tmap_bhld(&tv, v.ptr, v.batch_stride, v.row_stride, B, Lk, H, KB))) return rc;
  constexpr int NSLOT = MASKED ? 1 : MAX_BLOCKS_UNMASKED;
  const int smem = 1024 + Q_BYTES + 2 * NSLOT * KV_BYTES + 2 * P_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<MASKED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("attn_fwd_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  FwdArgs a;
  a.o = (__nv_bfloat16*)o.ptr; a.o_bs = o.batch_stride; a.o_rs = o.row_stride;
  a.lse = lse; a.tt = tt; a.Lq = Lq; a.Lk = Lk; a.H = H; a.n = n; a.Ti = Ti;
  a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Lq + TQ - 1) / TQ, H, B);
  attn_fwd_tc_kernel<MASKED><<<grid, FWD_THREADS, smem, st>>>(tq, tk, tv, a);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

int launch_attn_fwd_tc(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt, unimp_mview_t o,
                       float* lse, int B, int Lq, int Lk, int H, int n, int Ti, float scale,
                       cudaStream_t st) {
  if (tt) return launch_fwd<true>(q, k, v, tt, o, lse, B, Lq, Lk, H, n, Ti, scale, st);
  return launch_fwd<false>(q, k, v, nullptr, o, lse, B, Lq, Lk, H, n, Ti, scale, st);
}

bool attn_bwd_tc_supported(unimp_view_t, unimp_view_t, unimp_view_t, unimp_view_t, unimp_mview_t,
                           unimp_mview_t, unimp_mview_t, const int32_t*, int, int, int, int) {
  return false;
}
int launch_attn_bwd_tc(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*, unimp_view_t, unimp_view_t,
                       const float*, void*, unimp_mview_t, unimp_mview_t, unimp_mview_t, int, int, int, int,
                       int, int, float, cudaStream_t) {
  set_error("attn_bwd_tc: not built");
  return UNIMP_E_SHAPE;
}

}  // namespace unimp
