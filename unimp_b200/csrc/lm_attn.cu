// K4: causal self-attention of the (frozen) GPT-NeoX decoder layers, bf16, head dim 80, on tcgen05.
//
// Replaces the attention core of HF `GPTNeoXAttention.forward` (transformers gpt_neox; reached from
// upstream `FlamingoLayer.forward` -> `decoder_layer(...)`, call site UniMP/mmrec.py:177-181): after
// rotary, softmax(scale * q k^T + causal/padding mask) v per head.  SURVEY.md §8 row f3.
//
// Head dim 80 = TWO column panels read from tensor maps whose innermost extent is 80: a 64-column
// SWIZZLE_128B panel and a 16-column SWIZZLE_32B panel (coordinate 64).  S = Q K^T takes 4 K-steps
// from panel 0 and ONE from panel 1 (K = 80 exactly); O and the gradients are produced per panel
// (N = 64 and N = 16).  No padded copy of q/k/v exists in memory.
//
// Mask: key j is visible to query i iff j <= i and key_bits[b][j] (bit j%32 of word j/32; NULL = all
// keys valid) — what HF builds from a 2-D attention_mask for a causal LM.  Rows that see no key
// get o = 0 / lse = -inf.
//
// Forward : flash_fwd.cu (one-sweep online softmax, shared with the ViT / Perceiver cores).
// Backward: this file.  A first version (one CTA-wide step at a time, per-thread fp32 vector atomics
//           for dQ) measured 917 us at T=1024; the pipelined kernel below 352 us
//           (profiles/r2_lm_attn_*.log).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace unimp {
int make_tmap_tiled_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
namespace lm {

using namespace tc;

constexpr int TQ = 128, KB = 64, DH = 80;
constexpr uint32_t QP = TQ * 128;     // one [128 rows][64 cols] bf16 panel, 16 KB
constexpr uint32_t KP = KB * 128;     // one [64 rows][64 cols] panel, 8 KB
constexpr uint32_t PB = TQ * 128;     // P / dS: [128 rows][64 keys], 16 KB

struct Args {
  unsigned long long* dbg;   // test hook (unimp__lm_bwd_debug): 64 x u64 stamps per CTA, NULL = off
  __nv_bfloat16* o;          // fwd: out (B,T,H*80); bwd: forward output (read)
  const __nv_bfloat16* d_o;  // bwd
  int64_t o_bs, o_rs;        // shared by o and d_o (both (B,T,H*80) contiguous views)
  float* lse;                // (B,H,T)
  const float* delta;        // bwd: (B,H,T) rowsum(dO o O), written by lm_delta_kernel
  const uint32_t* kbits;     // (B, kwords) or NULL
  int kwords;
  float* dq32;               // bwd: (B,H,Tp,84) fp32 (Tp = T rounded up to 128), zeroed by the launcher
  __nv_bfloat16 *dk, *dv;    // bwd: (B,T,H,80) bf16 contiguous
  int T, H, B;
  int flags;                 // experiments (UNIMP_LM_BWD_FLAGS): 1 = skip the dQ bulk reduce, 2 = skip dQ staging too, 4 = head-major items
  float scale, scale_log2;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// n (multiple of 8) fp32 TMEM words -> bf16 row, scaled
__device__ __forceinline__ void store_bf16(__nv_bfloat16* dst, const uint32_t* r, int n, float mul) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c * 8 < n) {
      uint4 v;
      v.x = pack2(__uint_as_float(r[8 * c + 0]) * mul, __uint_as_float(r[8 * c + 1]) * mul);
      v.y = pack2(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul);
      v.z = pack2(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul);
      v.w = pack2(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul);
      *reinterpret_cast<uint4*>(dst + c * 8) = v;
    }
  }
}

// Which of the 64 keys of block j may query `row` of sample b see?
__device__ __forceinline__ uint64_t visible(const Args& a, int b, int j, int row) {
  const int d = row - j * KB;                 // last causal column
  uint64_t m = d < 0 ? 0ull : (d >= 63 ? ~0ull : ((2ull << d) - 1ull));
  if (a.kbits) {
    const uint32_t* w = a.kbits + (int64_t)b * a.kwords + 2 * j;
    m &= (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// Backward, pipelined.  CTA = (64-key block j, head, sample) walks the query tiles i at or below the
// diagonal.  Per pair, with lse and delta = rowsum(dO o O) precomputed per row:
//     S  = Q_i K_j^T            dP = dO_i V_j^T                          (128 x 64, K = 80)
//     P  = exp(scale*S - lse),  dS = scale * P o (dP - delta)            (threads)
//     [dV_j | dK_j] += [P | dS]^T [dO_i | Q_i]  per column panel (M = 128; N = 128 and N = 32)
//     dQ_i  = dS K_j             per panel; staged as fp32 [128][84] and added into dq32 with ONE
//                                asynchronous bulk reduce (cp.reduce.async.bulk .add.f32, 43 KB)
// Roles (320 threads): warps 0-7 = two threads per query row, each owns 32 of the 64 key columns;
// warp 8 = one elected lane issues every tcgen05.mma; warp 9 = one elected lane issues every TMA.
// Pipeline: Q/dO tiles travel through a 3-stage ring (full / empty mbarriers); S and dP are
// double-buffered in TMEM, so S/dP of pair i+1 run while the threads work on pair i and the
// gradient MMAs of pair i run while they stage dQ of pair i-1.
// TMEM (512 columns): S 2x64 | dP 2x64 | dKV p0 128 | dQ p0 64 | dKV p1 32 | dQ p1 16.
// ---------------------------------------------------------------------------------------------
// stamps of a CTA's FIRST item: 0 start (ns), 1 ready (ns), 2 last pair's gradients seen (ns), 3 item end
// (ns), 4 ready (cycles), 5 flush start (cycles), 6 item end (cycles), 7 CTA end (ns); pair i < 6 ->
// 8 + i*8 + k (cycles): 0 MMA lane got P_i,
// 1 MMA lane done with pair i, 2 worker S/dP ready, 3 math done, 4 previous gradients seen,
// 5 arrived, 6 dQ of the previous pair staged
#define BW_STAMP(slot)                                                                                   \
  do {                                                                                                   \
    if (a.dbg) {                                                                                         \
      unsigned long long t__;                                                                            \
      if ((slot) < 4 || (slot) == 7) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));              \
      else t__ = (unsigned long long)clock64();                                                          \
      a.dbg[blockIdx.x * 64 + (slot)] = t__;                                                             \
    }                                                                                                    \
  } while (0)

constexpr int NQS = 3, BW_THREADS = 2 * TQ + 64, BW_MMA = 8, BW_TMA = 9;
constexpr uint32_t QP1 = TQ * 32, KP1 = KB * 32;                   // 16-column SWIZZLE_32B panels
constexpr uint32_t ST_DO0 = 0, ST_Q0 = QP, ST_DO1 = 2 * QP, ST_Q1 = 2 * QP + QP1, ST_BYTES = 2 * QP + 2 * QP1;
constexpr uint32_t B_RING = 0, B_P = NQS * ST_BYTES, B_DS = B_P + PB, B_K0 = B_DS + PB, B_K1 = B_K0 + KP,
                   B_V0 = B_K1 + KP1, B_V1 = B_V0 + KP, B_STG = B_V1 + KP1, DQ_PITCH = 84,
                   STG_BYTES = TQ * DQ_PITCH * 4, B_BAR = B_STG + STG_BYTES, B_SMEM = B_BAR + 256;
static_assert(B_K1 % 1024 == 0 && B_V0 % 1024 == 0 && B_V1 % 1024 == 0 && B_STG % 1024 == 0, "tile alignment");

__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               :
               : "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// delta[b][h][t] = sum_d o[b][t][h][d] * d_o[b][t][h][d]   (one thread per row, t fastest)
__global__ void __launch_bounds__(256) lm_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                       const __nv_bfloat16* __restrict__ d_o,
                                                       float* __restrict__ delta, int T, int H, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = (int)(i % T);
  const int64_t bh = i / T;
  const int h = (int)(bh % H);
  const int64_t b = bh / H;
  const int64_t off = ((b * T + t) * H + h) * DH;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    Vec16<__nv_bfloat16> ov, gv;
    float of[8], gf[8];
    ov.load(o + off + c * 8); gv.load(d_o + off + c * 8);
    ov.unpack(of); gv.unpack(gf);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(of[e], gf[e], acc);
  }
  delta[i] = acc;
}

__global__ void __launch_bounds__(BW_THREADS, 1)
lm_attn_bwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tdo,
                   const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                   const __grid_constant__ CUtensorMap tq1, const __grid_constant__ CUtensorMap tdo1,
                   const __grid_constant__ CUtensorMap tk1, const __grid_constant__ CUtensorMap tv1,
                   const Args a) {
  // PERSISTENT: one CTA per SM walks work items (key block, head, sample), heavy key blocks first.
  // TMEM, the barriers and the three roles live across items; the producer fetches the next item's
  // K/V and first Q/dO tiles while the threads flush the current item's dK/dV (a 1-2 pair item at
  // T=256 spent 2/3 of its life in prologue, flush and CTA turnover: profiles/r2_lm_attn_bwd_timeline.log).
  constexpr uint32_t S_COL = 0, DP_COL = 128, DKV0_COL = 256, DQ0_COL = 384, DKV1_COL = 448, DQ1_COL = 480,
                     TMEM_COLS = 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_BAR);
  uint64_t *kv_full = bars, *kv_free = bars + 1, *qdo_full = bars + 2, *qdo_free = qdo_full + NQS,
           *bar_s = qdo_free + NQS, *bar_p = bar_s + 2, *bar_g = bar_s + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = tid < 2 * TQ;
  if (tid == 0) BW_STAMP(0);
  const int nq = (a.T + TQ - 1) / TQ, nkb = (a.T + KB - 1) / KB;
  const int HB = a.H * a.B, n_items = nkb * HB;
  // item w -> key block w / (H*B) (small key blocks = many query tiles first), head, sample.
  // (a.flags & 4: head-major order instead — the 16 key blocks of a head run side by side and its Q / dO
  // tiles stay in L2, 604 -> ~250 MB of DRAM reads at T = 1024 — measured -3 % there but +10 % at T = 256,
  // where the heavy-first order balances the CTAs: DRAM is not what bounds this kernel.)
  auto decode = [&](int w, int& kb, int& h, int& b) {
    int rem;
    if (a.flags & 4) {
      rem = w / nkb;
      kb = w - rem * nkb;
    } else {
      kb = w / HB;
      rem = w - kb * HB;
    }
    b = rem / a.H;
    h = rem - b * a.H;
  };

  if (warp == BW_MMA) {
    if (elect_one_sync()) {
      if (smem_u32(smem) & 1023u) {
        printf("unimp: lm_attn_bwd: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
      }
      mbar_init(kv_full, 1); mbar_init(kv_free, 1); mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1);
      mbar_init(bar_p, 8); mbar_init(bar_g, 1);
#pragma unroll
      for (int i = 0; i < NQS; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_free[i], 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t su = smem_u32(smem);
  if (tid == 0) { BW_STAMP(1); BW_STAMP(4); }

  // g = running pair count of this CTA (all roles count alike): Q/dO ring stage g % 3 (fill g / 3),
  // S/dP buffer g & 1 (use g / 2), bar_p / bar_g phase g; it = running item count (kv_full / kv_free).
  if (warp == BW_TMA && elect_one_sync()) {
    int it = 0, st = 0;
    uint32_t g = 0, free_par = 1;                          // parity of the stage's previous release
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int kb, h, b;
      decode(w, kb, h, b);
      const int q_lo = (kb * KB) / TQ, n = nq - q_lo;
      if (it > 0) mbar_wait(kv_free, (it - 1) & 1);        // the previous item's last MMAs have read K / V
      mbar_arrive_expect_tx(kv_full, 2 * KP + 2 * KP1);
      tma_load_4d(smem + B_K0, &tk, kv_full, 0, h, kb * KB, b);
      tma_load_4d(smem + B_K1, &tk1, kv_full, 64, h, kb * KB, b);
      tma_load_4d(smem + B_V0, &tv, kv_full, 0, h, kb * KB, b);
      tma_load_4d(smem + B_V1, &tv1, kv_full, 64, h, kb * KB, b);
#pragma unroll 1
      for (int i = 0; i < n; ++i, ++g) {
        uint8_t* d = smem + B_RING + st * ST_BYTES;
        const int r0 = (q_lo + i) * TQ;
        if (g >= NQS) mbar_wait(&qdo_free[st], free_par);
        mbar_arrive_expect_tx(&qdo_full[st], ST_BYTES);
        tma_load_4d(d + ST_Q0, &tq, &qdo_full[st], 0, h, r0, b);
        tma_load_4d(d + ST_Q1, &tq1, &qdo_full[st], 64, h, r0, b);
        tma_load_4d(d + ST_DO0, &tdo, &qdo_full[st], 0, h, r0, b);
        tma_load_4d(d + ST_DO1, &tdo1, &qdo_full[st], 64, h, r0, b);
        if (st == NQS - 1) { st = 0; free_par ^= 1; } else { ++st; }
      }
    }
  } else if (warp == BW_MMA && elect_one_sync()) {
    constexpr uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);         // S, dP
    constexpr uint32_t idesc_dq0 = make_idesc(TQ, 64, 0, 1);       // dQ = dS K   (B = K panel, MN-major)
    constexpr uint32_t idesc_dq1 = make_idesc(TQ, 16, 0, 1);
    constexpr uint32_t idesc_dkv0 = make_idesc(128, 128, 1, 1);    // [P|dS]^T [dO|Q] panel 0
    constexpr uint32_t idesc_dkv1 = make_idesc(128, 32, 1, 1);     // panel 1: N = 16 + 16
    // Descriptors of the fixed tiles, built once and kept as (lo, hi) halves: K-steps and ring stages
    // are 32-bit adds of (byte offset >> 4) to the start-address field in `lo` (every tile stays below
    // the 256 KB the field covers); ring stage and parities are carried incrementally, no div / mod.
    auto lo = [](uint64_t d) { return (uint32_t)d; };
    auto hi = [](uint64_t d) { return (uint32_t)(d >> 32); };
    const uint64_t D_k0 = make_smem_desc(su + B_K0, 16, 1024), D_k1 = make_smem_desc32(su + B_K1, 16, 256);
    const uint64_t D_v0 = make_smem_desc(su + B_V0, 16, 1024), D_v1 = make_smem_desc32(su + B_V1, 16, 256);
    const uint64_t D_k0mn = make_smem_desc(su + B_K0, 1024, 1024), D_k1mn = make_smem_desc32(su + B_K1, 256, 256);
    const uint64_t D_p = make_smem_desc(su + B_P, PB, 1024), D_ds = make_smem_desc(su + B_DS, 16, 1024);
    const uint64_t D_q0 = make_smem_desc(su + B_RING + ST_Q0, 16, 1024),
                   D_q1 = make_smem_desc32(su + B_RING + ST_Q1, 16, 256);
    const uint64_t D_do0 = make_smem_desc(su + B_RING + ST_DO0, 16, 1024),
                   D_do1 = make_smem_desc32(su + B_RING + ST_DO1, 16, 256);
    const uint64_t D_do0mn = make_smem_desc(su + B_RING + ST_DO0, QP, 1024),
                   D_do1mn = make_smem_desc32(su + B_RING + ST_DO1, QP1, 256);
    const uint32_t h128 = hi(D_k0), h32 = hi(D_k1);       // hi halves: SBO 1024 / SWIZZLE_128B, SBO 256 / SWIZZLE_32B
    constexpr uint32_t ST16 = ST_BYTES >> 4;
    // S / dP of a pair into S/dP buffer `buf`, from ring stage `st` (so = st * ST16), fill parity `par`
    auto issue_sdp = [&](int st, uint32_t so, int buf, uint32_t par) {
      mbar_wait(&qdo_full[st], par);
      tcgen05_fence_after();
      const uint32_t q0l = lo(D_q0) + so, do0l = lo(D_do0) + so;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss_lohi(tmem + S_COL + buf * 64, q0l + k4 * 2, h128, lo(D_k0) + k4 * 2, h128, idesc_s, k4 > 0);
      umma_ss_lohi(tmem + S_COL + buf * 64, lo(D_q1) + so, h32, lo(D_k1), h32, idesc_s, 1);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss_lohi(tmem + DP_COL + buf * 64, do0l + k4 * 2, h128, lo(D_v0) + k4 * 2, h128, idesc_s, k4 > 0);
      umma_ss_lohi(tmem + DP_COL + buf * 64, lo(D_do1) + so, h32, lo(D_v1), h32, idesc_s, 1);
      umma_commit(&bar_s[buf]);
    };
    // running state of pair g: ring stage / its fill parity, and the same for pairs g+1, g+2 (look-ahead)
    int it = 0;
    uint32_t g = 0;
    int st0 = 0, st1 = 1, st2 = 2;                 // stages of pairs g, g+1, g+2
    uint32_t par0 = 0, par1 = 0, par2 = 0;         // fill parities of those stages for those pairs
    auto advance = [&]() {                         // g -> g + 1
      ++g;
      st0 = st1; par0 = par1;
      st1 = st2; par1 = par2;
      if (st2 == NQS - 1) { st2 = 0; par2 ^= 1; } else { ++st2; }
    };
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int kb, h, b;
      decode(w, kb, h, b);
      const int n = nq - (kb * KB) / TQ;
      mbar_wait(kv_full, it & 1);
      issue_sdp(st0, st0 * ST16, g & 1, par0);
      if (n > 1) issue_sdp(st1, st1 * ST16, (g + 1) & 1, par1);
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        const uint32_t so = st0 * ST16;
        mbar_wait(bar_p, g & 1);            // P / dS in shared memory, S/dP buffer and dQ columns released
        if (it == 0 && i < 6) BW_STAMP(8 + i * 8 + 0);
        tcgen05_fence_after();
        const uint32_t do0mn = lo(D_do0mn) + so, do1mn = lo(D_do1mn) + so;
        const uint32_t hp = hi(D_p), hdo0 = hi(D_do0mn), hdo1 = hi(D_do1mn), hds = hi(D_ds), hk0 = hi(D_k0mn),
                       hk1 = hi(D_k1mn);
#pragma unroll
        for (int k8 = 0; k8 < TQ / 16; ++k8)
          umma_ss_lohi(tmem + DKV0_COL, lo(D_p) + k8 * 128, hp, do0mn + k8 * 128, hdo0, idesc_dkv0, (i > 0 || k8 > 0));
#pragma unroll
        for (int k8 = 0; k8 < TQ / 16; ++k8)
          umma_ss_lohi(tmem + DKV1_COL, lo(D_p) + k8 * 128, hp, do1mn + k8 * 32, hdo1, idesc_dkv1, (i > 0 || k8 > 0));
#pragma unroll
        for (int k4 = 0; k4 < KB / 16; ++k4)
          umma_ss_lohi(tmem + DQ0_COL, lo(D_ds) + k4 * 2, hds, lo(D_k0mn) + k4 * 128, hk0, idesc_dq0, k4 > 0);
#pragma unroll
        for (int k4 = 0; k4 < KB / 16; ++k4)
          umma_ss_lohi(tmem + DQ1_COL, lo(D_ds) + k4 * 2, hds, lo(D_k1mn) + k4 * 32, hk1, idesc_dq1, k4 > 0);
        umma_commit(bar_g);
        umma_commit(&qdo_free[st0]);
        if (i + 1 == n) umma_commit(kv_free);
        if (i + 2 < n) issue_sdp(st2, st2 * ST16, g & 1, par2);
        if (it == 0 && i < 6) BW_STAMP(8 + i * 8 + 1);
        advance();
      }
    }
  }

  if (worker) {
    const int r = tid & (TQ - 1), hs = tid >> 7;          // row in the tile, 32-column half
    const float l2e = 1.4426950408889634f;
    float* stg = reinterpret_cast<float*>(smem + B_STG);
    const int Tp = nq * TQ;
    if (hs == 1) *reinterpret_cast<float4*>(stg + r * DQ_PITCH + DH) = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t q0[32], q1[16];                              // dQ of the previous pair, on its way to staging
    int g = 0, it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int kb, h, b;
      decode(w, kb, h, b);
      const int q_lo = (kb * KB) / TQ, n = nq - q_lo;
      float* dq_head = a.dq32 + ((int64_t)b * a.H + h) * Tp * DQ_PITCH;
      const float* lse_h = a.lse + ((int64_t)b * a.H + h) * a.T;
      const float* del_h = a.delta + ((int64_t)b * a.H + h) * a.T;

      // dQ tile `qt` (already in q0 / q1) -> staging -> one bulk reduce into dq32
      auto stage_dq = [&](int qt) {
        if (a.flags & 2) return;
        // (bulk-copy instructions take uniform operands: one ELECTED lane of warp 0 issues, commits and
        // waits — under `tid == 0` ptxas wraps the reduce in a waterfall loop)
        if (warp == 0 && elect_one_sync()) tma_store_wait_read();   // the previous reduce has left the staging buffer
        workers_sync();
        float* dst = stg + r * DQ_PITCH + hs * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(dst + c * 4) =
              make_float4(__uint_as_float(q0[4 * c]), __uint_as_float(q0[4 * c + 1]), __uint_as_float(q0[4 * c + 2]),
                          __uint_as_float(q0[4 * c + 3]));
        if (hs == 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<float4*>(stg + r * DQ_PITCH + 64 + c * 4) =
                make_float4(__uint_as_float(q1[4 * c]), __uint_as_float(q1[4 * c + 1]), __uint_as_float(q1[4 * c + 2]),
                            __uint_as_float(q1[4 * c + 3]));
        }
        fence_proxy_async_smem();
        workers_sync();
        if (warp == 0 && elect_one_sync() && !(a.flags & 1)) {
          bulk_reduce_add_f32(dq_head + (int64_t)qt * TQ * DQ_PITCH, stg, STG_BYTES);
          tma_store_commit();
        }
      };
      auto read_dq = [&]() {
        tmem_ld32(lane_addr + DQ0_COL + hs * 32, q0);
        if (hs == 1) tmem_ld16(lane_addr + DQ1_COL, q1);
        tmem_ld_wait();
      };

#pragma unroll 1
      for (int i = 0; i < n; ++i, ++g) {
        const int row = (q_lo + i) * TQ + r;
        const bool valid = row < a.T;
        float lse_l2 = 0.f, delta = 0.f;
        uint32_t vh = 0u;
        if (valid) {
          const float lse = lse_h[row];
          delta = del_h[row];
          lse_l2 = lse * l2e;
          if (lse > -INFINITY) vh = (uint32_t)(visible(a, b, kb, row) >> (32 * hs));
        }
        mbar_wait(&bar_s[g & 1], (g >> 1) & 1);
        tcgen05_fence_after();
        if (tid == 0 && it == 0 && i < 6) BW_STAMP(8 + i * 8 + 2);
        uint32_t sv[32], dp[32];
        tmem_ld32(lane_addr + S_COL + (g & 1) * 64 + hs * 32, sv);
        tmem_ld32(lane_addr + DP_COL + (g & 1) * 64 + hs * 32, dp);
        tmem_ld_wait();
        uint4 pp[4], dd[4];
        uint32_t* pw = reinterpret_cast<uint32_t*>(pp);
        uint32_t* dw = reinterpret_cast<uint32_t*>(dd);
        const float nd = -delta;
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float p0 = exp2f(fmaf(__uint_as_float(sv[c]), a.scale_log2, -lse_l2));
          float p1 = exp2f(fmaf(__uint_as_float(sv[c + 1]), a.scale_log2, -lse_l2));
          if (vh != 0xffffffffu) {
            if (!(vh & (1u << c))) p0 = 0.f;
            if (!(vh & (2u << c))) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dp[c]) + nd) * a.scale;
          const float d1 = p1 * (__uint_as_float(dp[c + 1]) + nd) * a.scale;
          pw[c >> 1] = pack2(p0, p1);
          dw[c >> 1] = pack2(d0, d1);
        }
        if (tid == 0 && it == 0 && i < 6) BW_STAMP(8 + i * 8 + 3);
        if (i > 0) {                     // gradient MMAs of pair i-1 done: sP / sdS free, dQ_{i-1} readable
          mbar_wait(bar_g, (g - 1) & 1);
          tcgen05_fence_after();
        }
        if (tid == 0 && it == 0 && i < 6) BW_STAMP(8 + i * 8 + 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(smem + B_P + sw128_offset(r, hs * 4 + c)) = pp[c];
          *reinterpret_cast<uint4*>(smem + B_DS + sw128_offset(r, hs * 4 + c)) = dd[c];
        }
        fence_proxy_async_smem();
        if (i > 0) read_dq();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        if (tid == 0 && it == 0 && i < 6) BW_STAMP(8 + i * 8 + 5);
        if (i > 0) stage_dq(q_lo + i - 1);   // overlaps the gradient MMAs of pair i
        if (tid == 0 && it == 0 && i < 6) BW_STAMP(8 + i * 8 + 6);
      }
      mbar_wait(bar_g, (g - 1) & 1);
      tcgen05_fence_after();
      if (tid == 0 && it == 0) { BW_STAMP(2); BW_STAMP(5); }
      read_dq();
      stage_dq(nq - 1);

      // ---- flush dV (lanes 0-63) and dK (lanes 64-127) of this key block; each thread 32 (+16) columns
      {
        const int key = kb * KB + (r & 63);
        const bool is_k = r >= 64;
        __nv_bfloat16* dst = (is_k ? a.dk : a.dv) + (((int64_t)b * a.T + key) * a.H + h) * DH;
        tmem_ld32(lane_addr + DKV0_COL + (is_k ? 64 : 0) + hs * 32, q0);
        if (hs == 1) tmem_ld16(lane_addr + DKV1_COL + (is_k ? 16 : 0), q1);
        tmem_ld_wait();
        if (key < a.T) {
          store_bf16(dst + hs * 32, q0, 32, 1.f);
          if (hs == 1) store_bf16(dst + 64, q1, 16, 1.f);
        }
      }
      tcgen05_fence_before();     // orders the TMEM reads above before this warp's next bar_p arrival
      if (tid == 0 && it == 0) { BW_STAMP(3); BW_STAMP(6); }
    }
    if (warp == 0 && elect_one_sync()) tma_store_wait_read();   // the last reduce has left shared memory (its
                                                                // global side completes before the grid does)
    if (tid == 0) BW_STAMP(7);
  }
  __syncthreads();
  if (warp == BW_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

// attention_mask (B,T) (nonzero = real token) -> one bit per key, 32 keys per word
template <typename M>
__global__ void key_bits_kernel(const M* __restrict__ mask, uint32_t* __restrict__ bits, int T, int kwords) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int w = threadIdx.x >> 5; w < kwords; w += nw) {
    const int key = w * 32 + lane;
    const bool on = key < T && mask[(int64_t)b * T + key] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[(int64_t)b * kwords + w] = word;
  }
}

// ---- host -----------------------------------------------------------------------------------
// panel 0: 64 columns SWIZZLE_128B from column 0; panel 1: 16 columns SWIZZLE_32B (coordinate 64)
static int make_map_p(CUtensorMap* out, const void* base, int64_t bs, int64_t rs, int64_t hs, int B, int T,
                      int H, int box_rows, int p1) {
  if (B == 1) bs = rs * (int64_t)T;
  const uint64_t dims[4] = {(uint64_t)DH, (uint64_t)H, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)hs * 2, (uint64_t)rs * 2, (uint64_t)bs * 2};
  const uint32_t box[4] = {p1 ? 16u : 64u, 1, (uint32_t)box_rows, 1};
  return make_tmap_tiled_sw(out, base, 4, dims, strides, box, p1 ? 32 : 128);
}

static unsigned long long* g_bwd_dbg = nullptr;

static const char* unsupported(const void* q, const void* k, const void* v, int64_t bs, int64_t rs,
                               int64_t hs, int dh, int dtype) {
  if (dtype != UNIMP_BF16) return "bf16 only";
  if (dh != DH) return "head dim must be 80";
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || bs % 8 || rs % 8 || hs % 8)
    return "q/k/v must be 16-byte aligned views with strides that are multiples of 8 elements";
  return nullptr;
}

}  // namespace lm
}  // namespace unimp

namespace unimp {
int launch_flash_fwd_80(const void* q, const void* k, const void* v, int64_t bs, int64_t rs, int64_t hs,
                        const uint32_t* kbits, void* o, float* lse, int B, int T, int H, float scale,
                        cudaStream_t st);
}

using namespace unimp;

extern "C" int unimp_lm_attn_supported(int T, int H, int dh, int dtype) {
  return dtype == UNIMP_BF16 && dh == lm::DH && T >= 1 && H >= 1 && H <= 65535;
}

extern "C" int unimp_key_bits(const void* mask, int elem_size, uint32_t* bits, int B, int T, void* stream) {
  UNIMP_CHECK_ARG(mask && bits, UNIMP_E_NULL, "unimp_key_bits: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0, UNIMP_E_SHAPE, "unimp_key_bits: empty shape");
  UNIMP_CHECK_ARG(elem_size == 1 || elem_size == 8, UNIMP_E_DTYPE, "unimp_key_bits: mask must be bool/uint8 or int64");
  const int kwords = 2 * ((T + 63) / 64);
  if (elem_size == 8)
    lm::key_bits_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const int64_t*)mask, bits, T, kwords);
  else
    lm::key_bits_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)mask, bits, T, kwords);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_lm_attn_fwd(const void* q, const void* k, const void* v, int64_t batch_stride,
                                 int64_t row_stride, int64_t head_stride, const uint32_t* key_bits,
                                 void* o, float* lse, int B, int T, int H, int dh, float scale, int dtype,
                                 void* stream) {
  UNIMP_CHECK_ARG(q && k && v && o && lse, UNIMP_E_NULL, "unimp_lm_attn_fwd: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0 && H > 0, UNIMP_E_SHAPE, "unimp_lm_attn_fwd: empty shape");
  const char* why = lm::unsupported(q, k, v, batch_stride, row_stride, head_stride, dh, dtype);
  UNIMP_CHECK_ARG(!why, UNIMP_E_SHAPE, "unimp_lm_attn_fwd: %s", why);
  UNIMP_CHECK_ARG(aligned16(o), UNIMP_E_ALIGN, "unimp_lm_attn_fwd: o must be 16-byte aligned");
  return launch_flash_fwd_80(q, k, v, batch_stride, row_stride, head_stride, key_bits, o, lse, B, T, H, scale,
                             (cudaStream_t)stream);
}

extern "C" int unimp_lm_attn_bwd(const void* q, const void* k, const void* v, int64_t batch_stride,
                                 int64_t row_stride, int64_t head_stride, const uint32_t* key_bits,
                                 const void* o, const void* d_o, const float* lse, float* dq32, float* delta,
                                 void* dk, void* dv, int B, int T, int H, int dh, float scale, int dtype,
                                 void* stream) {
  UNIMP_CHECK_ARG(q && k && v && o && d_o && lse && dq32 && delta && dk && dv, UNIMP_E_NULL,
                  "unimp_lm_attn_bwd: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0 && H > 0, UNIMP_E_SHAPE, "unimp_lm_attn_bwd: empty shape");
  const char* why = lm::unsupported(q, k, v, batch_stride, row_stride, head_stride, dh, dtype);
  UNIMP_CHECK_ARG(!why, UNIMP_E_SHAPE, "unimp_lm_attn_bwd: %s", why);
  UNIMP_CHECK_ARG(aligned16(o) && aligned16(d_o) && aligned16(dq32) && aligned16(dk) && aligned16(dv),
                  UNIMP_E_ALIGN, "unimp_lm_attn_bwd: o/d_o/dq32/dk/dv must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap m[8];
  int rc;
  const int64_t D = (int64_t)H * dh;
  for (int p1 = 0; p1 < 2; ++p1) {
    if ((rc = lm::make_map_p(&m[4 * p1 + 0], q, batch_stride, row_stride, head_stride, B, T, H, lm::TQ, p1))) return rc;
    if ((rc = lm::make_map_p(&m[4 * p1 + 1], d_o, (int64_t)T * D, D, dh, B, T, H, lm::TQ, p1))) return rc;
    if ((rc = lm::make_map_p(&m[4 * p1 + 2], k, batch_stride, row_stride, head_stride, B, T, H, lm::KB, p1))) return rc;
    if ((rc = lm::make_map_p(&m[4 * p1 + 3], v, batch_stride, row_stride, head_stride, B, T, H, lm::KB, p1))) return rc;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lm::lm_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)lm::B_SMEM);
    if (e != cudaSuccess) { set_error("lm_attn_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  const int Tp = (T + lm::TQ - 1) / lm::TQ * lm::TQ;
  cudaError_t e = cudaMemsetAsync(dq32, 0, (size_t)B * H * Tp * lm::DQ_PITCH * sizeof(float), st);
  if (e != cudaSuccess) { set_error("lm_attn_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
  const int64_t rows = (int64_t)B * H * T;
  lm::lm_delta_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)o,
                                                                      (const __nv_bfloat16*)d_o, delta, T, H, rows);
  UNIMP_CHECK_LAUNCH();
  lm::Args a{};
  a.o = (__nv_bfloat16*)const_cast<void*>(o); a.d_o = (const __nv_bfloat16*)d_o;
  a.o_bs = (int64_t)T * D; a.o_rs = D;
  a.lse = const_cast<float*>(lse); a.delta = delta; a.kbits = key_bits; a.kwords = 2 * ((T + 63) / 64);
  a.dq32 = dq32; a.dk = (__nv_bfloat16*)dk; a.dv = (__nv_bfloat16*)dv;
  a.dbg = lm::g_bwd_dbg;
  static const int flags = getenv("UNIMP_LM_BWD_FLAGS") ? atoi(getenv("UNIMP_LM_BWD_FLAGS")) : 0;
  a.flags = flags;
  a.T = T; a.H = H; a.B = B; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  const int64_t n_items = (int64_t)((T + lm::KB - 1) / lm::KB) * H * B;
  static int n_sms = 0;
  if (!n_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sms <= 0)
      n_sms = UNIMP_NUM_SMS;
  }
  const unsigned grid = (unsigned)(n_items < n_sms ? n_items : n_sms);   // persistent: one CTA per SM
  lm::lm_attn_bwd_kernel<<<grid, lm::BW_THREADS, lm::B_SMEM, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], a);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

// Test hook (not in the public header): 64 x u64 stamps per CTA of the next unimp_lm_attn_bwd launches.
extern "C" void unimp__lm_bwd_debug(unsigned long long* buf) { unimp::lm::g_bwd_dbg = buf; }
