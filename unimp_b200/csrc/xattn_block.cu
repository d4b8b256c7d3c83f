// K1-fused — MaskedCrossAttention as ONE kernel: to_q GEMM -> masked media-located attention ->
// to_out GEMM (upstream `MaskedCrossAttention.forward` after its LayerNorm; call site
// reference UniMP/mmrec.py:177-181).  ~8.4 GFLOP per launch at configs[1] instead of 0.2 for the
// bare attention core: this is the variant whose roofline is the tensor pipe.
//
//   y (B,T,D) = softmax_mask( (x_ln Wq^T) K^T * scale ) V  Wout^T        H = 8 heads x 64
//
// One thread-block CLUSTER of 8 CTAs per 128-row query tile, one head per CTA:
//   phase 1  Q_h (128x64, TMEM) = x_ln tile (128 x D) . Wq_h^T, K = D in 64-wide chunks through a
//            ring of 3 stages x 2 chunks (144 KB in flight per SM).  Two chunks per stage because the
//            MMA issuer pays ~250 cycles per mbarrier wait (measured): one wait per 8 MMAs, not 4.
//            The x_ln chunk is the same for all 8 heads: CTA (k mod 8) loads chunk k ONCE and
//            TMA-MULTICASTS it into all 8 CTAs' rings (L2 reads of x_ln / 8); a ring slot is
//            released by tcgen05.commit multicast to all 8 CTAs' "empty" barriers, because any of
//            them may be the one that overwrites it next.
//   phase 2  the attention core of attn_tc.cu (S = Q K^T, mask from text_time in registers, row
//            softmax on the 128 TMEM-lane-owning threads, O = P V); Q goes TMEM -> bf16 ->
//            swizzled shared memory (A operand) and to global (saved for backward).
//   phase 3  O_h (bf16) goes to global (saved for backward); after a cluster barrier every CTA
//            multicasts ITS head's O tile back from L2 into all 8 CTAs (one mbarrier per tile: the MMAs
//            of K chunk kk start when tile kk is in), so each holds the full 128 x 512 O tile as an A
//            operand, and computes 1/8 of the output columns:
//            y[:, c*D/8 : (c+1)*D/8] = O_all . Wout[c*D/8 : ...]^T  (Wout streamed by TMA, 2 stages).
// Warp roles (192 threads): warps 0-3 own one query row / TMEM lane each (conversion, softmax,
// epilogues); warp 4 = TMA producer, warp 5 = MMA issuer (warp 5 owns TMEM): one lane of each,
// chosen by elect.sync so that the uniform-register instructions go out without waterfall loops.
// Every mbarrier wait is bounded and tagged (a protocol bug traps and names the barrier).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace unimp {

using namespace tc;

namespace xb {
constexpr int TQ = 128, KB = 64, DH = 64, H = 8, INNER = H * DH;
constexpr int THREADS = 192;
constexpr int S1 = 3;                                  // phase-1 ring: 3 stages of TWO K chunks (48 KB) each; the
                                                       // chunk slots 3-5 live in the idle Wout ring
constexpr uint32_t A_BYTES = TQ * 64 * 2;              // 16 KB: x_ln chunk (128 rows x 64)
constexpr uint32_t B_BYTES = 64 * 64 * 2;              // 8 KB: Wq_h chunk
constexpr uint32_t STAGE1 = A_BYTES + B_BYTES;         // 24 KB
constexpr uint32_t Q_BYTES = TQ * DH * 2, P_BYTES = TQ * KB * 2, KV_BYTES = KB * DH * 2;
constexpr int NS_MAX = 320;                            // output columns per CTA (D / 8), D <= 2560
constexpr uint32_t W_STAGE = NS_MAX * 128;             // 40 KB: Wout rows [c*NS, +NS) x 64 K-columns
constexpr uint32_t OFF_RING1 = 0;
constexpr uint32_t OFF_Q = OFF_RING1 + 3 * STAGE1;     //  73728
constexpr uint32_t OFF_P = OFF_Q + Q_BYTES;            //  90112
constexpr uint32_t OFF_K = OFF_P + P_BYTES;            // 106496
constexpr uint32_t OFF_V = OFF_K + 2 * KV_BYTES;       // 122880
constexpr uint32_t OFF_W = OFF_V + 2 * KV_BYTES;       // 139264
constexpr uint32_t OFF_O = 0;                          // phase 3: 8 x 16 KB over the dead ring / Q / P / K / V
constexpr uint32_t SMEM_BYTES = OFF_W + 2 * W_STAGE;   // 221184
static_assert(OFF_O + H * Q_BYTES <= OFF_W, "the O tiles must not reach the Wout ring");
static_assert(3 * STAGE1 <= 2 * W_STAGE, "ring stages 3-5 live in the Wout ring while it is idle");
// phase-1 ring stage i: 0-2 at the front, 3-5 in the Wout ring (not needed before phase 2)
__device__ __forceinline__ uint32_t stage1_off(int i) {
  return i < 3 ? OFF_RING1 + (uint32_t)i * STAGE1 : OFF_W + (uint32_t)(i - 3) * STAGE1;
}
constexpr uint32_t Q_COL = 0, S_COL = 64, O_COL = 128, Y_COL = 0, TMEM_COLS = 512;
constexpr uint16_t ALL = 0xFF;

enum Tag { T_FULL1 = 10, T_EMPTY1 = 20, T_QACC = 30, T_QS = 31, T_KV = 40, T_S = 50, T_P = 51, T_PV = 52,
           T_O = 53, T_FULLO = 60, T_FULLW = 70, T_EMPTYW = 80, T_Y = 90 };
}  // namespace xb

struct XBlkArgs {
  __nv_bfloat16 *q, *o, *y;   // q, o: (B,T,512) contiguous (saved for backward); y: (B,T,D) contiguous
  float* lse;                 // (B,H,T)
  const int32_t* tt;          // (B,T)
  int T, Ti, n, D, NK, NS, NSH, nsplit;
  float scale, scale_log2;
  unsigned long long* dbg;    // optional per-CTA phase timestamps (16 x u64 per CTA), NULL = off
  int flags;                  // bit 0 = no multicast: every CTA loads its own x_ln chunks and releases its
                              // ring locally — the DEFAULT since it measured 3 % faster (25.3 vs 26.1 us:
                              // multicast couples the eight CTAs' ring slots and L2 bandwidth is not the
                              // limit); UNIMP_XB_FLAGS=0 restores the multicast form for A/B runs
};

__device__ __forceinline__ unsigned long long xb_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define XB_STAMP(slot)                                                                          \
  do {                                                                                          \
    if (a.dbg) a.dbg[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = xb_now(); \
  } while (0)

__device__ __forceinline__ uint32_t xb_pack(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// 32 fp32 values -> 32 bf16 = 4 x 16-byte chunks (chunk index half*4 + c) of row `row` of a
// SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ void xb_store_half(uint8_t* tile, int row, int half, const uint32_t* r, float mul) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    v.x = xb_pack(__uint_as_float(r[8 * c + 0]) * mul, __uint_as_float(r[8 * c + 1]) * mul);
    v.y = xb_pack(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul);
    v.z = xb_pack(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul);
    v.w = xb_pack(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul);
    *reinterpret_cast<uint4*>(tile + sw128_offset(row, half * 4 + c)) = v;
  }
}
__device__ __forceinline__ void xb_store_global32(__nv_bfloat16* dst, const uint32_t* r, float mul, int ncols) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c * 8 < ncols) {
      uint4 v;
      v.x = xb_pack(__uint_as_float(r[8 * c + 0]) * mul, __uint_as_float(r[8 * c + 1]) * mul);
      v.y = xb_pack(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul);
      v.z = xb_pack(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul);
      v.w = xb_pack(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul);
      *reinterpret_cast<uint4*>(dst + c * 8) = v;
    }
  }
}

__global__ void __launch_bounds__(xb::THREADS, 1)
xattn_block_fwd_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap twq,
                       const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                       const __grid_constant__ CUtensorMap to, const __grid_constant__ CUtensorMap twout,
                       const __grid_constant__ CUtensorMap tqo, const __grid_constant__ CUtensorMap ty,
                       const XBlkArgs a) {
  using namespace xb;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + OFF_Q;
  uint8_t* sP = smem + OFF_P;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sW = smem + OFF_W;
  uint8_t* sO = smem + OFF_O;
  __shared__ uint64_t full1[S1], empty1[S1], bar_qacc, bar_qs, bar_kv[2], bar_s, bar_p, bar_pv, bar_o,
      full_o[xb::H], full_w[2], empty_w[2], bar_y;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_j[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = tid < TQ;
  const int h = (int)cluster_ctarank();          // head of phases 1-2, output-column slice of phase 3
  const int t0 = blockIdx.y * TQ, b = blockIdx.z;
  const int NK = a.NK, NS = a.NS, NSH = a.NSH;

  if (warp == 4) {
    if (elect_one_sync()) {
#pragma unroll
      for (int i = 0; i < S1; ++i) { mbar_init(&full1[i], 1); mbar_init(&empty1[i], (a.flags & 1) ? 1 : H); }
      mbar_init(&bar_qacc, 1); mbar_init(&bar_qs, 4);
      mbar_init(&bar_kv[0], 1); mbar_init(&bar_kv[1], 1);
      mbar_init(&bar_s, 1); mbar_init(&bar_p, 4); mbar_init(&bar_pv, 1); mbar_init(&bar_o, 1);
#pragma unroll
      for (int i = 0; i < H; ++i) mbar_init(&full_o[i], 1);
      mbar_init(&full_w[0], 1); mbar_init(&full_w[1], 1);
      mbar_init(&empty_w[0], 1); mbar_init(&empty_w[1], 1);
      mbar_init(&bar_y, 1);
      fence_barrier_init();
      tma_prefetch_desc(&tx); tma_prefetch_desc(&twq); tma_prefetch_desc(&twout);
      // per-CTA x_ln loads touch only THIS CTA's barriers: the first three ring stages go out now, under
      // the text_time reads, the TMEM allocation and the CTA-wide sync (phase 1 is paced by data arrival)
      if (a.flags & 1) {
#pragma unroll
        for (int s = 0; s < S1; ++s) {
          if (s < a.NK / 2) {
            mbar_arrive_expect_tx(&full1[s], 2 * STAGE1);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int k = 2 * s + c;
              uint8_t* st = smem + stage1_off(2 * s + c);
              tma_load_3d(st, &tx, &full1[s], k * 64, t0, b);
              tma_load_2d(st + A_BYTES, &twq, &full1[s], k * 64, h * DH);
            }
          }
        }
      }
    }
    // image blocks referenced by the tile's rows (same for the 8 CTAs of the cluster)
    int lo = 1 << 30, hi = -1;
#pragma unroll
    for (int i = 0; i < TQ / 32; ++i) {
      const int r = t0 + lane + 32 * i;
      if (r < a.T) {
        const int t = a.tt[(int64_t)b * a.T + r];
        if (t > a.Ti) { lo = 0; hi = a.Ti - 1; }
        else if (t >= 1) { lo = min(lo, t - 1); hi = max(hi, t - 1); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (elect_one_sync()) {
      s_j[0] = lo; s_j[1] = hi;
      // loads that touch only THIS CTA's barriers may start before the cluster is in step:
      // K/V of the first two image blocks
      const int nb = hi >= lo ? hi - lo + 1 : 0;
      for (int it = 0; it < nb && it < 2; ++it) {
        mbar_arrive_expect_tx(&bar_kv[it], 2 * KV_BYTES);
        tma_load_4d(sK + it * KV_BYTES, &tk, &bar_kv[it], 0, h, (lo + it) * a.n, b);
        tma_load_4d(sV + it * KV_BYTES, &tv, &bar_kv[it], 0, h, (lo + it) * a.n, b);
      }
    }
    __syncwarp();
  }
  if (warp == 5) tmem_alloc(&tmem_slot, TMEM_COLS);

  const int row = t0 + tid;
  const bool valid = worker && row < a.T;
  int ttr = 0;
  if (valid) ttr = a.tt[(int64_t)b * a.T + row];
  const bool uniform = ttr > a.Ti;
  const int blk = (ttr >= 1 && !uniform) ? ttr - 1 : -1;

  if (tid == 0) XB_STAMP(0);
  // barrier #1: every CTA's barriers are initialised before any multicast / remote arrive.  Without
  // multicast nothing crosses CTAs before cluster barrier #2: a CTA-wide barrier is enough.
  tcgen05_fence_before();
  if (a.flags & 1) __syncthreads();
  else cluster_sync_all();
  tcgen05_fence_after();
  if (tid == 0) XB_STAMP(1);
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int jlo = s_j[0], jhi = s_j[1];
  const int nblk = jhi >= jlo ? jhi - jlo + 1 : 0;

  const uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);      // Q-proj chunk, S = Q K^T
  const uint32_t idesc_o = make_idesc(TQ, DH, 0, 1);      // O = P V (V MN-major)
  const uint32_t idesc_y = make_idesc(TQ, NSH, 0, 0);     // y slice = O_all Wout^T

  // =============================== phases 1 + 2 =============================================
  if (warp == 4) {
    if (elect_one_sync()) {
      // one ring stage = TWO 64-wide K chunks (48 KB).  The stage loop is unrolled by the ring depth so
      // that every shared-memory offset is a compile-time constant: the elected lane runs on the
      // uniform datapath, where each extra address instruction costs ~12 cycles.
      const int NP = NK / 2;
      uint32_t ph = 1;                                // parity of the "empty" phase to wait for
      for (int p0 = 0; p0 < NP; p0 += S1) {
#pragma unroll
        for (int s = 0; s < S1; ++s) {
          const int p = p0 + s;
          if (p < NP && !(p0 == 0 && (a.flags & 1))) {      // (p0 == 0 went out in the prologue)
            if (p0 > 0) mbar_wait_tag(&empty1[s], ph, T_EMPTY1 + s);
            mbar_arrive_expect_tx(&full1[s], 2 * STAGE1);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int k = 2 * p + c;
              uint8_t* st = smem + stage1_off(2 * s + c);
              if (a.flags & 1) tma_load_3d(st, &tx, &full1[s], k * 64, t0, b);               // own copy of the chunk
              else if ((k & (H - 1)) == h) tma_load_3d_mc(st, &tx, &full1[s], k * 64, t0, b, ALL);
              tma_load_2d(st + A_BYTES, &twq, &full1[s], k * 64, h * DH);
            }
          }
        }
        ph ^= 1;
      }
      // the Wout ring doubles as ring stages 3-5: its first two chunks are fetched once every
      // phase-1 MMA of this CTA has completed (they arrive while phase 2 runs)
      mbar_wait_tag(&bar_qacc, 0, T_QACC);
      for (int kk = 0; kk < 2; ++kk) {
        mbar_arrive_expect_tx(&full_w[kk], (uint32_t)NS * 128u);
        for (int half = 0; half < a.nsplit; ++half)
          tma_load_2d(sW + kk * W_STAGE + half * NSH * 128, &twout, &full_w[kk], kk * 64, h * NS + half * NSH);
      }
      for (int it = 0; it + 2 < nblk; ++it) {     // K/V refills (tiles that span > 2 images)
        mbar_wait_tag(&bar_pv, it & 1, T_PV);
        const int st = it & 1;
        mbar_arrive_expect_tx(&bar_kv[st], 2 * KV_BYTES);
        tma_load_4d(sK + st * KV_BYTES, &tk, &bar_kv[st], 0, h, (jlo + it + 2) * a.n, b);
        tma_load_4d(sV + st * KV_BYTES, &tv, &bar_kv[st], 0, h, (jlo + it + 2) * a.n, b);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one_sync()) {
      const int NP = NK / 2;
      const uint32_t smem_base = smem_u32(smem);
      uint32_t ph = 0;
      for (int p0 = 0; p0 < NP; p0 += S1) {
#pragma unroll
        for (int s = 0; s < S1; ++s) {
          const int p = p0 + s;
          if (p < NP) {
            mbar_wait_tag(&full1[s], ph, T_FULL1 + s);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint32_t sa = smem_base + stage1_off(2 * s + c), sb = sa + A_BYTES;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_ss(tmem + Q_COL, make_smem_desc(sa + k4 * 32, 16, 1024), make_smem_desc(sb + k4 * 32, 16, 1024),
                        idesc_s, (p > 0 || c > 0 || k4 > 0));
            }
            if (a.flags & 1) umma_commit(&empty1[s]);   // CTA-local ring when nothing is multicast
            else umma_commit_mc(&empty1[s], ALL);       // the stage is free in THIS CTA; all 8 must say so
          }
        }
        ph ^= 1;
      }
      umma_commit(&bar_qacc);
      if (nblk > 0) {
        mbar_wait_tag(&bar_qs, 0, T_QS);
        mbar_wait_tag(&bar_kv[0], 0, T_KV);
        tcgen05_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < DH / 16; ++k4)
          umma_ss(tmem + S_COL, make_smem_desc(smem_u32(sQ) + k4 * 32, 16, 1024),
                  make_smem_desc(smem_u32(sK) + k4 * 32, 16, 1024), idesc_s, k4 > 0);
        umma_commit(&bar_s);
      }
      for (int it = 0; it < nblk; ++it) {
        const int st = it & 1;
        mbar_wait_tag(&bar_p, it & 1, T_P);
        tcgen05_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < KB / 16; ++k4)
          umma_ss(tmem + O_COL, make_smem_desc(smem_u32(sP) + k4 * 32, 16, 1024),
                  make_smem_desc(smem_u32(sV + st * KV_BYTES) + k4 * 2048, 1024, 1024), idesc_o,
                  (it > 0 || k4 > 0));
        umma_commit(&bar_pv);
        if (it + 1 == nblk) umma_commit(&bar_o);
        if (it + 1 < nblk) {
          const int sn = (it + 1) & 1;
          mbar_wait_tag(&bar_kv[sn], ((it + 1) >> 1) & 1, T_KV + 1);
          tcgen05_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < DH / 16; ++k4)
            umma_ss(tmem + S_COL, make_smem_desc(smem_u32(sQ) + k4 * 32, 16, 1024),
                    make_smem_desc(smem_u32(sK + sn * KV_BYTES) + k4 * 32, 16, 1024), idesc_s, k4 > 0);
          umma_commit(&bar_s);
        }
      }
    }
    __syncwarp();
  } else {
    // ---- workers: Q accumulators -> bf16 -> shared (A operand of S) + global (saved for backward)
    uint32_t r[32];
    mbar_wait_tag(&bar_qacc, 0, T_QACC);
    tcgen05_fence_after();
    if (tid == 0) XB_STAMP(2);     // phase 1 done
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tmem_ld32(lane_addr + Q_COL + half * 32, r);
      tmem_ld_wait();
      xb_store_half(sQ, tid, half, r, 1.f);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_qs);
    // q for backward: the swizzled tile just written IS the TMA box layout -> one bulk store instead of
    // 128 B of strided STG per thread (rows beyond T are clipped by the tensor map)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 0 && elect_one_sync()) {
      tma_store_4d(&tqo, sQ, 0, h, t0, b);
      tma_store_commit();
    }

    // ---- masked row softmax per image block (as xattn_fwd_tc_kernel) ------------------------
    float sum = 0.f, m_row = 0.f;
    for (int it = 0; it < nblk; ++it) {
      const int j = jlo + it;
      mbar_wait_tag(&bar_s, it & 1, T_S);
      tcgen05_fence_after();
      const bool mine = uniform || blk == j;
      const bool warp_any = __any_sync(0xffffffffu, mine);
      float m = -INFINITY;
      if (warp_any) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(lane_addr + S_COL + half * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) m = fmaxf(m, __uint_as_float(r[c]));
        }
      }
      if (uniform) m = 0.f;
      const float ms = m * a.scale_log2;
      float psum = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (warp_any) {
          tmem_ld32(lane_addr + S_COL + half * 32, r);
          tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float v = 0.f;
          if (mine) v = uniform ? 1.f : exp2f(__uint_as_float(r[c]) * a.scale_log2 - ms);
          psum += v;
          r[c] = __float_as_uint(v);
        }
        xb_store_half(sP, tid, half, r, 1.f);
      }
      if (mine) { sum += psum; m_row = m; }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p);
    }
    // ---- O / sum -> global (saved for backward; re-read by the cluster in phase 3) ----------
    const float lse_val = sum > 0.f ? m_row * a.scale + logf(sum) : -INFINITY;
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (nblk > 0) {
      mbar_wait_tag(&bar_o, 0, T_O);
      tcgen05_fence_after();
    }
    if (tid == 0) XB_STAMP(3);     // attention MMAs done
    // normalised O -> bf16 -> the (dead) P tile in the TMA box layout -> ONE bulk store (rows beyond T are
    // clipped by the tensor map) instead of 128 bytes of strided STG per thread
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (nblk > 0) {
        tmem_ld32(lane_addr + O_COL + half * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = 0u;
      }
      xb_store_half(sP, tid, half, r, inv);
    }
    if (valid) a.lse[((int64_t)b * H + h) * a.T + row] = lse_val;
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 0 && elect_one_sync()) {
      tma_store_4d(&to, sP, 0, h, t0, b);
      tma_store_commit();
      // complete (not only read): the peers' multicast loads after the cluster barrier fetch this tile
      // from L2, and they overwrite sQ / sP next — the q store is in the same wait
      tma_store_wait_all();
      fence_proxy_async_all();
    }
    tcgen05_fence_before();
  }

  // cluster barrier #2: every head's O tile is in global memory (L2), every CTA is done with its
  // ring / Q / P / K / V buffers and with the Q / S / O accumulator columns
  if (tid == 0) XB_STAMP(4);       // o stored
  cluster_sync_all();
  tcgen05_fence_after();
  if (tid == 0) XB_STAMP(5);       // cluster barrier #2 passed

  // =============================== phase 3: y slice = O_all . Wout_slice^T ====================
  if (warp == 4) {
    if (elect_one_sync()) {
      fence_proxy_async_all();
      // one barrier per head's tile: the to_out MMAs of K chunk kk start when tile kk is in, under the
      // remaining tiles' arrival
#pragma unroll
      for (int i = 0; i < H; ++i) mbar_arrive_expect_tx(&full_o[i], Q_BYTES);
      tma_load_4d_mc(sO + h * Q_BYTES, &to, &full_o[h], 0, h, t0, b, ALL);
      for (int kk = 2; kk < H; ++kk) {
        const int s = kk & 1;
        mbar_wait_tag(&empty_w[s], ((kk >> 1) - 1) & 1, T_EMPTYW + s);
        mbar_arrive_expect_tx(&full_w[s], (uint32_t)NS * 128u);
        for (int half = 0; half < a.nsplit; ++half)
          tma_load_2d(sW + s * W_STAGE + half * NSH * 128, &twout, &full_w[s], kk * 64, h * NS + half * NSH);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one_sync()) {
      const uint32_t so_base = smem_u32(sO), sw_base = smem_u32(sW);
      const uint32_t half_off = (uint32_t)NSH * 128u;
#pragma unroll
      for (int kk = 0; kk < H; ++kk) {              // fully unrolled: constant offsets, constant parities
        const int s = kk & 1;
        mbar_wait_tag(&full_o[kk], 0, T_FULLO + kk);
        if (kk == 0) XB_STAMP(6);  // the first O tile arrived
        mbar_wait_tag(&full_w[s], (kk >> 1) & 1, T_FULLW + s);
        tcgen05_fence_after();
        const uint32_t sa = so_base + kk * Q_BYTES, sb = sw_base + s * W_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          umma_ss(tmem + Y_COL, make_smem_desc(sa + k4 * 32, 16, 1024), make_smem_desc(sb + k4 * 32, 16, 1024),
                  idesc_y, (kk > 0 || k4 > 0));
          if (a.nsplit > 1)
            umma_ss(tmem + Y_COL + NSH, make_smem_desc(sa + k4 * 32, 16, 1024),
                    make_smem_desc(sb + half_off + k4 * 32, 16, 1024), idesc_y, (kk > 0 || k4 > 0));
        }
        umma_commit(&empty_w[s]);
      }
      umma_commit(&bar_y);
    }
    __syncwarp();
  } else {
    uint32_t r[32];
    mbar_wait_tag(&bar_y, 0, T_Y);
    tcgen05_fence_after();
    if (tid == 0) XB_STAMP(7);     // to_out MMAs done
    if (NS % 64 == 0) {
      // stage the slice as 64-column swizzled tiles over the (now dead) O tiles and let TMA write
      // them: full 128-byte lines instead of 40 strided 16-byte stores per thread
      uint8_t* sY = smem + OFF_O;
      for (int c64 = 0; c64 < NS / 64; ++c64) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(lane_addr + Y_COL + c64 * 64 + half * 32, r);
          tmem_ld_wait();
          xb_store_half(sY + c64 * Q_BYTES, tid, half, r, 1.f);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 0 && elect_one_sync()) {
          tma_store_3d(&ty, sY + c64 * Q_BYTES, h * NS + c64 * 64, t0, b);
          tma_store_commit();
        }
      }
      if (warp == 0 && elect_one_sync()) tma_store_wait_read();
    } else {
      __nv_bfloat16* yrow = a.y + ((int64_t)b * a.T + row) * a.D + h * NS;
      for (int c = 0; c < NS; c += 32) {
        tmem_ld32(lane_addr + Y_COL + c, r);
        tmem_ld_wait();
        if (valid) xb_store_global32(yrow + c, r, 1.f, NS - c < 32 ? NS - c : 32);
      }
    }
    tcgen05_fence_before();
  }
  if (tid == 0) XB_STAMP(8);       // y stored
  // nobody leaves while a peer may still write into its shared memory.  Without x_ln multicast the only
  // remote writes are the O tiles, and every CTA has waited for all eight of them (full_o): a CTA-wide
  // barrier (all tcgen05.ld done before TMEM is freed) is enough.
  if (a.flags & 1) __syncthreads();
  else cluster_sync_all();
  if (tid == 0) XB_STAMP(9);
  if (warp == 5) tmem_dealloc(tmem, TMEM_COLS);
}

// ---- host -------------------------------------------------------------------------------------

static unsigned long long* g_xb_dbg = nullptr;   // test hook: per-CTA phase timestamps

const char* xattn_block_unsupported(int T, int Ti, int n, int Hh, int dh, int D, int dtype) {
  (void)T; (void)Ti;
  if (dtype != UNIMP_BF16) return "bf16 only (fp32 runs the unfused path)";
  if (Hh != xb::H || dh != xb::DH) return "heads must be 8 x 64 (one head per CTA of an 8-CTA cluster)";
  if (n != xb::KB) return "n_latents must be 64";
  if (D % 128 != 0 || D > 8 * xb::NS_MAX) return "D must be a multiple of 128 and <= 2560";
  const int NS = D / 8;
  if (NS > 256 && (NS / 2) % 16 != 0) return "D/16 must be a multiple of 16 when D/8 > 256";
  return nullptr;
}

int launch_xattn_block_fwd(const void* x_ln, const void* w_q, unimp_view_t k, unimp_view_t v,
                           const int32_t* tt, const void* w_out, void* q, void* o, float* lse, void* y,
                           int B, int T, int Ti, int n, int D, float scale, cudaStream_t st) {
  using namespace xb;
  CUtensorMap tx, twq, tk, tv, to, twout, tqo, ty;
  int rc;
  const int NS = D / 8, nsplit = NS > 256 ? 2 : 1, NSH = NS / nsplit;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)T * D * 2};
    const uint32_t box[3] = {64, (uint32_t)TQ, 1};
    if ((rc = make_tmap_tiled(&tx, x_ln, 3, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)INNER};
    const uint64_t strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap_tiled(&twq, w_q, 2, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)INNER, (uint64_t)D};
    const uint64_t strides[1] = {(uint64_t)INNER * 2};
    const uint32_t box[2] = {64, (uint32_t)NSH};
    if ((rc = make_tmap_tiled(&twout, w_out, 2, dims, strides, box))) return rc;
  }
  if ((rc = make_tmap_bhld(&tk, k.ptr, k.batch_stride, k.row_stride, B, Ti * n, H, KB))) return rc;
  if ((rc = make_tmap_bhld(&tv, v.ptr, v.batch_stride, v.row_stride, B, Ti * n, H, KB))) return rc;
  if ((rc = make_tmap_bhld(&to, o, (int64_t)T * INNER, INNER, B, T, H, TQ))) return rc;
  if ((rc = make_tmap_bhld(&tqo, q, (int64_t)T * INNER, INNER, B, T, H, TQ))) return rc;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)T * D * 2};
    const uint32_t box[3] = {64, (uint32_t)TQ, 1};
    if ((rc = make_tmap_tiled(&ty, y, 3, dims, strides, box))) return rc;
  }
  const int smem = 1024 + SMEM_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(xattn_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("xattn_block_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  XBlkArgs a;
  a.q = (__nv_bfloat16*)q; a.o = (__nv_bfloat16*)o; a.y = (__nv_bfloat16*)y; a.lse = lse; a.tt = tt;
  a.T = T; a.Ti = Ti; a.n = n; a.D = D; a.NK = D / 64; a.NS = NS; a.NSH = NSH; a.nsplit = nsplit;
  a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.dbg = g_xb_dbg;
  {
    static int flags = -1;
    if (flags < 0) { const char* e = getenv("UNIMP_XB_FLAGS"); flags = e ? atoi(e) : 1; }
    a.flags = flags;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H, (T + TQ - 1) / TQ, B);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = H; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, xattn_block_fwd_kernel, tx, twq, tk, tv, to, twout, tqo, ty, a);
  if (e != cudaSuccess) { set_error("xattn_block_fwd launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // namespace unimp

// Test hook (not in the public header): device buffer of 16 x u64 per CTA that the next launches
// fill with %globaltimer phase stamps; NULL switches it off.
extern "C" void unimp__xattn_block_debug(unsigned long long* buf) { unimp::g_xb_dbg = buf; }
