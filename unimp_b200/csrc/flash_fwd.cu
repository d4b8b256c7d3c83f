// One-sweep attention forward on tcgen05 (round 2, second generation of the unmasked / causal cores).
//
//   K2  Perceiver latent attention 64 x 320, K3 ViT-L/14 self-attention 257 x 257  (DH = 64)
//   K4  GPT-NeoX causal self-attention with key padding                            (DH = 80)
//
// Common to both kernels below (vs round 2's first, two-sweep kernels: S recomputed, 3 CTAs per SM):
//   * ONE sweep over the key blocks with an online softmax; O is rescaled in TMEM only when a row's
//     running maximum grows by more than 2^8 (rare after the first blocks), so the common step is
//     S -> registers -> exp2 -> P (bf16) with no TMEM round trip of O;
//   * P reaches the PV MMA through TENSOR memory: the row's thread writes its 64 probabilities with
//     tcgen05.st (32 columns of packed bf16 pairs) and the MMA reads its A operand from TMEM — no
//     shared-memory store, no generic->async proxy fence, PV reads only V from shared memory;
//   * K and V travel through two 3-stage TMA rings fed by a dedicated producer lane (full / empty
//     mbarriers: tcgen05.commit releases a stage);
//   * head dim 80 = a 64-column SWIZZLE_128B panel + a 16-column SWIZZLE_32B panel (2 KB per 64-row
//     tile instead of the 8 KB a zero-padded 128-byte panel costs) read through two tensor maps over
//     the same memory.
// flash_fwd_kernel : CTA = (128-query tile, head, sample), two CTAs per SM; S double-buffered in TMEM
//     (S_{j+1} = Q K_{j+1}^T runs while the 128 softmax threads work on S_j); warps 0-3: one query row
//     per thread (TMEM lane = row), warp 4: MMA lane, warp 5: TMA lane.
// flash_fwd3_kernel: three query tiles per CTA, one persistent CTA per SM (see its header); chosen by
//     launch_flash for large workloads.
// All mbarrier waits are bounded.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace unimp {
namespace ff {

using namespace tc;

constexpr int TQ = 128, KB = 64, THREADS = TQ + 64;   // 4 softmax warps + MMA warp + TMA warp
constexpr int W_MMA = 4, W_TMA = 5;
constexpr int NST = 3;                        // K and V ring depth
constexpr uint32_t QP0 = TQ * 128;            // [128][64] bf16, SWIZZLE_128B
constexpr uint32_t QP1 = TQ * 32;             // [128][16] bf16, SWIZZLE_32B
constexpr uint32_t KP0 = KB * 128, KP1 = KB * 32;
constexpr uint32_t PB = TQ * 128;

struct Args {
  unsigned long long* dbg;   // test hook (unimp__flash_fwd_debug): 64 x u64 %globaltimer stamps per CTA, NULL = off
  __nv_bfloat16* o;
  int64_t o_bs, o_rs;        // o[b][row][h*DH + d]
  float* lse;                // (B,H,Lq)
  const uint32_t* kbits;     // (B, kwords) key-valid bits or NULL
  int kwords;
  int Lq, Lk, H, Bt;
  float scale, scale_log2;
};

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// stamp slot: step j < 8 -> 8 + j*6 + k ; k: 0 issuer got P_j, 1 issuer done with step j,
// 2 worker got S_j, 3 worker exps done, 4 worker got PV_{j-1}, 5 worker arrived
#define FF_STAMP(slot)                                                                                   \
  do {                                                                                                   \
    if (a.dbg) a.dbg[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 + (slot)] =            \
        (slot) < 4 ? now_ns() : (unsigned long long)clock64();                                          \
  } while (0)

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
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

template <bool CAUSAL>
__device__ __forceinline__ uint64_t visible(const Args& a, int b, int j, int row) {
  uint64_t m = ~0ull;
  if (CAUSAL) {
    const int d = row - j * KB;
    m = d < 0 ? 0ull : (d >= 63 ? ~0ull : ((2ull << d) - 1ull));
  }
  const int rem = a.Lk - j * KB;               // keys of this block that exist
  if (rem < KB) m &= rem <= 0 ? 0ull : ((1ull << rem) - 1ull);
  if (a.kbits) {
    const uint32_t* w = a.kbits + (int64_t)b * a.kwords + 2 * j;
    m &= (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  }
  return m;
}

template <int DH>
struct Smem {
  static constexpr bool P1 = DH > 64;
  static constexpr uint32_t KST = KP0 + (P1 ? KP1 : 0);       // one K (or V) stage
  static constexpr uint32_t Q0 = 0, Q1 = QP0, KR = Q1 + (P1 ? QP1 : 0), VR = KR + NST * KST,
                            P = VR + NST * KST, BAR = P + PB, TOTAL = BAR + 256;
};

template <int DH, bool CAUSAL, bool PT>
__global__ void __launch_bounds__(THREADS, 2)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                 const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tq1,
                 const __grid_constant__ CUtensorMap tk1, const __grid_constant__ CUtensorMap tv1, const Args a) {
  using L = Smem<DH>;
  constexpr bool P1 = L::P1;
  // PT: P goes to the MMA through TENSOR memory (tcgen05.st by the row's thread, A operand of PV read
  // from TMEM) instead of shared memory: no 16 KB store + generic->async proxy fence per step, and the PV
  // MMAs read only V from shared memory.
  constexpr uint32_t S_COL = 0, O0_COL = 128, O1_COL = 192, P_COL = 208, TMEM_COLS = 256;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  // full barriers (TMA -> MMA), empty barriers (tcgen05.commit -> TMA producer), S / P / PV handshakes
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bar_k + NST, *k_free = bar_v + NST, *v_free = k_free + NST,
           *bar_s = v_free + NST, *bar_p = bar_s + 2, *bar_pv = bar_s + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = tid < TQ;
  if (tid == 0) FF_STAMP(0);
  const int qt = CAUSAL ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;   // causal: long tiles first
  const int row0 = qt * TQ, h = blockIdx.y, b = blockIdx.z;
  const int nb_all = (a.Lk + KB - 1) / KB;
  const int nb = CAUSAL ? min(nb_all, (row0 + TQ) / KB) : nb_all;

  if (warp == W_MMA) {
    if (elect_one_sync()) {
      if (smem_u32(smem) & 1023u) {
        printf("unimp: flash_fwd: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
      }
      mbar_init(bar_q, 1); mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1); mbar_init(bar_p, 4);
      mbar_init(bar_pv, 1);
#pragma unroll
      for (int i = 0; i < NST; ++i) {
        mbar_init(&bar_k[i], 1); mbar_init(&bar_v[i], 1); mbar_init(&k_free[i], 1); mbar_init(&v_free[i], 1);
      }
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
  if (tid == 0) { FF_STAMP(1); FF_STAMP(4); }

  // The two single-lane roles run mostly on the uniform datapath (~12 cycles per dependent
  // instruction): their loops are unrolled by 6 = lcm(ring depth 3, 2 S buffers) so that every
  // stage offset, S buffer and most barrier parities are compile-time constants.
  if (warp == W_TMA && elect_one_sync()) {
    // ---- TMA producer: Q, then K_j / V_j as their ring stages are released -----------------
    mbar_arrive_expect_tx(bar_q, QP0 + (P1 ? QP1 : 0));
    tma_load_4d(smem + L::Q0, &tq, bar_q, 0, h, row0, b);
    if (P1) tma_load_4d(smem + L::Q1, &tq1, bar_q, 64, h, row0, b);
    for (int j0 = 0; j0 < nb; j0 += 6) {
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int j = j0 + u;
        if (j < nb) {
          const int st = u % NST;                                  // compile-time after unrolling
          const uint32_t free_par = ((u / NST) + 1) & 1;           // parity of fill (j / NST) - 1
          uint8_t* kd = smem + L::KR + st * L::KST;
          uint8_t* vd = smem + L::VR + st * L::KST;
          if (j >= NST) mbar_wait(&k_free[st], free_par);
          mbar_arrive_expect_tx(&bar_k[st], L::KST);
          tma_load_4d(kd, &tk, &bar_k[st], 0, h, j * KB, b);
          if (P1) tma_load_4d(kd + KP0, &tk1, &bar_k[st], 64, h, j * KB, b);
          if (j >= NST) mbar_wait(&v_free[st], free_par);
          mbar_arrive_expect_tx(&bar_v[st], L::KST);
          tma_load_4d(vd, &tv, &bar_v[st], 0, h, j * KB, b);
          if (P1) tma_load_4d(vd + KP0, &tv1, &bar_v[st], 64, h, j * KB, b);
        }
      }
    }
  } else if (warp == W_MMA && elect_one_sync()) {
    // ---- MMA issuer -------------------------------------------------------------------------
    constexpr uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);
    constexpr uint32_t idesc_o0 = make_idesc(TQ, 64, 0, 1);
    constexpr uint32_t idesc_o1 = make_idesc(TQ, 16, 0, 1);
    const uint32_t su = smem_u32(smem);
    // S_j = Q K_j^T into S buffer (j & 1); st = j % NST, par = parity of that stage's fill
    auto issue_s = [&](int st, int buf, uint32_t par) {
      const uint32_t k_u = su + L::KR + st * L::KST;
      const uint32_t d = tmem + S_COL + buf * 64;
      mbar_wait(&bar_k[st], par);
      tcgen05_fence_after();
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss(d, make_smem_desc(su + L::Q0 + k4 * 32, 16, 1024), make_smem_desc(k_u + k4 * 32, 16, 1024),
                idesc_s, k4 > 0);
      if (P1) umma_ss(d, make_smem_desc32(su + L::Q1, 16, 256), make_smem_desc32(k_u + KP0, 16, 256), idesc_s, 1);
      umma_commit(&bar_s[buf]);
      umma_commit(&k_free[st]);
    };
    mbar_wait(bar_q, 0);
    issue_s(0, 0, 0);
    if (nb > 1) issue_s(1, 1, 0);
    for (int j0 = 0; j0 < nb; j0 += 6) {
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int j = j0 + u;
        if (j < nb) {
          const int st = u % NST;
          mbar_wait(bar_p, u & 1);                       // P_j in shared memory, S_j consumed
          if (j < 8) FF_STAMP(8 + j * 6 + 0);
          mbar_wait(&bar_v[st], (u / NST) & 1);
          tcgen05_fence_after();
          const uint32_t v_u = su + L::VR + st * L::KST;
          if (PT) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_ts(tmem + O0_COL, tmem + P_COL + k4 * 8, make_smem_desc(v_u + k4 * 2048, 1024, 1024), idesc_o0,
                      (j > 0 || k4 > 0));
            if (P1) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_ts(tmem + O1_COL, tmem + P_COL + k4 * 8, make_smem_desc32(v_u + KP0 + k4 * 512, 256, 256),
                        idesc_o1, (j > 0 || k4 > 0));
            }
          } else {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_ss(tmem + O0_COL, make_smem_desc(su + L::P + k4 * 32, 16, 1024),
                      make_smem_desc(v_u + k4 * 2048, 1024, 1024), idesc_o0, (j > 0 || k4 > 0));
            if (P1) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_ss(tmem + O1_COL, make_smem_desc(su + L::P + k4 * 32, 16, 1024),
                        make_smem_desc32(v_u + KP0 + k4 * 512, 256, 256), idesc_o1, (j > 0 || k4 > 0));
            }
          }
          umma_commit(bar_pv);
          umma_commit(&v_free[st]);
          // S_{j+2} into the buffer the workers just released; stage (u+2) % 3, fill (j+2) / 3
          if (j + 2 < nb) issue_s((u + 2) % NST, u & 1, ((u + 2) / NST) & 1);
          if (j < 8) FF_STAMP(8 + j * 6 + 1);
        }
      }
    }
  }

  if (worker) {
    // One thread per query row (TMEM lane = row).  A second thread per row on the other 32 columns
    // was tried: it needs the row maximum over all 64 scores (a second TMEM read and max pass per
    // row) and the exponentials are MUFU-bound however many warps share them (535 cycles per
    // 128 x 64 block, tools/probes/softmax_probe.cu): 104 -> 109 us at T=1024
    // (profiles/r2_flash_fwd_timeline_two_threads_per_row.log).
    const int row = row0 + tid;
    const bool valid = row < a.Lq;
    const bool active = row0 + (warp << 5) < a.Lq;   // warp-uniform: any valid row in this warp
    float m_ref = -INFINITY, sum = 0.f;              // m_ref: the maximum the stored exponents refer to (raw S units)
    const float tau = 8.f / a.scale_log2;            // rescale O only when the maximum grows by > 2^8
    const float sl2 = a.scale_log2;
    float f0[32], f1[32];
    uint32_t* s0 = reinterpret_cast<uint32_t*>(f0);
    uint32_t* s1 = reinterpret_cast<uint32_t*>(f1);
    for (int j = 0; j < nb; ++j) {
      const uint64_t vis = (active && valid) ? visible<CAUSAL>(a, b, j, row) : 0ull;
      const bool any = __any_sync(0xffffffffu, vis != 0ull);
      mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
      tcgen05_fence_after();
      if (tid == 0 && j < 8) FF_STAMP(8 + j * 6 + 2);
      float alpha = 1.f;
      bool grow = false;
      uint4 pk[8];                                   // this row's 64 probabilities, bf16
      if (any) {
        const uint32_t sc = lane_addr + S_COL + (j & 1) * 64;
        tmem_ld32(sc, s0);
        tmem_ld32(sc + 32, s1);
        tmem_ld_wait();
        if (vis != ~0ull) {                          // diagonal / ragged / padded block: hide the masked keys
          const uint32_t v0 = (uint32_t)vis, v1 = (uint32_t)(vis >> 32);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (!(v0 & (1u << c))) f0[c] = -INFINITY;
            if (!(v1 & (1u << c))) f1[c] = -INFINITY;
          }
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 32; ++c) mx[c & 3] = fmaxf(mx[c & 3], fmaxf(f0[c], f1[c]));
        const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (bm > m_ref + tau || (m_ref == -INFINITY && bm > -INFINITY)) {
          grow = m_ref > -INFINITY;                  // O holds something to rescale
          alpha = grow ? exp2f((m_ref - bm) * sl2) : 1.f;
          m_ref = bm;
        }
        const float nms = m_ref > -INFINITY ? -m_ref * sl2 : 0.f;
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float e0 = exp2f(fmaf(f0[c], sl2, nms)), e1 = exp2f(fmaf(f0[c + 1], sl2, nms));
          const float g0 = exp2f(fmaf(f1[c], sl2, nms)), g1 = exp2f(fmaf(f1[c + 1], sl2, nms));
          ps[(c >> 1) & 3] += (e0 + e1) + (g0 + g1);
          pw[c >> 1] = pack2(e0, e1);
          pw[16 + (c >> 1)] = pack2(g0, g1);
        }
        sum = sum * alpha + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) pk[c] = make_uint4(0u, 0u, 0u, 0u);
      }
      // sP and O are free once PV_{j-1} has finished
      if (tid == 0 && j < 8) FF_STAMP(8 + j * 6 + 3);
      if (j > 0) {
        mbar_wait(bar_pv, (j - 1) & 1);
        tcgen05_fence_after();
        if (tid == 0 && j < 8) FF_STAMP(8 + j * 6 + 4);
        if (__any_sync(0xffffffffu, grow)) {
          uint32_t t[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld32(lane_addr + O0_COL + half * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) t[c] = __float_as_uint(__uint_as_float(t[c]) * alpha);
            tmem_st32(lane_addr + O0_COL + half * 32, t);
          }
          if (P1) {
            tmem_ld16(lane_addr + O1_COL, t);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) t[c] = __float_as_uint(__uint_as_float(t[c]) * alpha);
            tmem_st16(lane_addr + O1_COL, t);
          }
          tmem_st_wait();
        }
      }
      if (PT) {
        if (active) tmem_st32(lane_addr + P_COL, reinterpret_cast<const uint32_t*>(pk));
        tmem_st_wait();
      } else {
        if (active) {
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(smem + L::P + sw128_offset(tid, c)) = pk[c];
        }
        fence_proxy_async_smem();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (tid == 0 && j < 8) FF_STAMP(8 + j * 6 + 5);
    }
    mbar_wait(bar_pv, (nb - 1) & 1);
    tcgen05_fence_after();
    if (tid == 0) FF_STAMP(2);
    if (active) {
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      __nv_bfloat16* orow = a.o + (int64_t)b * a.o_bs + (int64_t)row * a.o_rs + h * DH;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld32(lane_addr + O0_COL + half * 32, s0);
        tmem_ld_wait();
        if (valid) store_bf16(orow + half * 32, s0, 32, inv);
      }
      if (P1) {
        tmem_ld16(lane_addr + O1_COL, s0);
        tmem_ld_wait();
        if (valid) store_bf16(orow + 64, s0, 16, inv);
      }
      if (valid) a.lse[((int64_t)b * a.H + h) * a.Lq + row] = sum > 0.f ? m_ref * a.scale + logf(sum) : -INFINITY;
    }
    tcgen05_fence_before();
    if (tid == 0) FF_STAMP(3);
  }
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Three query tiles per CTA (384 rows, 12 softmax warps), ONE CTA per SM, one K/V stream.
// The serial part of a softmax step (wait for S, TMEM read, wait for PV, store P, proxy fence, arrive,
// wait for the next S: ~1200 of the ~2350 cycles of a step, profiles/r2_flash_fwd_timeline.log) can only
// be hidden by OTHER warps; TMEM (2 x 256 columns) caps the two-CTA form at 8 softmax warps per SM.  Here
// S is single-buffered (64 columns per tile) so that three tiles fit (3 x (64 + 80) = 432 columns): while
// one tile's threads are in their serial part the other two keep the TMEM port and the MUFU busy, and
// every K/V block is fetched once for 384 query rows instead of once per 128.
// Causal: the tiles of a group have 6c+2, 6c+4, 6c+6 key blocks; a tile drops out when it is done.
// ---------------------------------------------------------------------------------------------
constexpr int NT = 3, THREADS3 = NT * TQ + 64, W3_MMA = NT * 4, W3_TMA = NT * 4 + 1;

template <int DH>
struct Smem3 {
  static constexpr bool P1 = DH > 64;
  static constexpr uint32_t QT = QP0 + (P1 ? QP1 : 0);        // one Q tile: panel 0 (+ panel 1)
  static constexpr uint32_t KST = KP0 + (P1 ? KP1 : 0);
  static constexpr uint32_t Q = 0, KR = NT * QT, VR = KR + NST * KST, P = VR + NST * KST, BAR = P + NT * PB,
                            TOTAL = BAR + 256;
};

template <int DH, bool CAUSAL, bool PT>
__global__ void __launch_bounds__(THREADS3, 1)
flash_fwd3_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                  const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tq1,
                  const __grid_constant__ CUtensorMap tk1, const __grid_constant__ CUtensorMap tv1, const Args a) {
  // PERSISTENT: gridDim.x CTAs (one per SM) walk work items (tile group, head, sample), long groups
  // first.  TMEM, barriers and roles live across items; the K/V rings run on into the next item and its Q
  // tiles are fetched as soon as this item's last S has been issued, so prologue and epilogue of
  // neighbouring items overlap (they were ~30 % of a ViT CTA's life: profiles/r2_flash_fwd3_timeline_per_cta.log).
  using L = Smem3<DH>;
  constexpr bool P1 = L::P1;
  constexpr uint32_t S_COL = 0, O0_COL = 64 * NT, O1_COL = 128 * NT, TMEM_COLS = 512;   // per tile: +64t, +64t, +16t
  // PT: P_t goes to the PV MMA through tensor memory.  Head dim 64 has room for it (columns 384 + 32t);
  // head dim 80 does not (3 x (64 + 80 + 32) > 512): there P_t overwrites the first 32 columns of S_t once
  // the row's thread has read them, and the MMA lane issues PV_t(j) BEFORE S_t(j+1) (the tensor pipe runs
  // in issue order, so S_t(j+1) cannot overtake the PV that still reads P_t).
  constexpr bool P_ALIAS = PT && P1;
  constexpr uint32_t P_COL = P_ALIAS ? S_COL : 128 * NT, P_STRIDE = P_ALIAS ? 64 : 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t *bar_q = bars, *q_free = bars + 1, *bar_k = bars + 2, *bar_v = bar_k + NST, *k_free = bar_v + NST,
           *v_free = k_free + NST, *bar_s = v_free + NST, *bar_p = bar_s + NT, *bar_pv = bar_p + NT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_pv + NT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) FF_STAMP(0);
  const int n_tiles = (a.Lq + TQ - 1) / TQ, n_grp = (n_tiles + NT - 1) / NT;
  const int HB = a.H * a.Bt, n_items = n_grp * HB;
  const int nb_all = (a.Lk + KB - 1) / KB;
  // item w -> group (causal: the long groups first), head, sample
  auto decode = [&](int w, int& grp, int& h, int& b) {
    const int gi = w / HB, rem = w - gi * HB;
    grp = CAUSAL ? n_grp - 1 - gi : gi;
    b = rem / a.H;
    h = rem - b * a.H;
  };
  auto nb_of = [&](int grp, int t) { return CAUSAL ? min(nb_all, ((grp * NT + t) * TQ + TQ) / KB) : nb_all; };

  if (warp == W3_MMA) {
    if (elect_one_sync()) {
      if (smem_u32(smem) & 1023u) {
        printf("unimp: flash_fwd3: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
      }
      mbar_init(bar_q, 1); mbar_init(q_free, 1);
#pragma unroll
      for (int i = 0; i < NST; ++i) {
        mbar_init(&bar_k[i], 1); mbar_init(&bar_v[i], 1); mbar_init(&k_free[i], 1); mbar_init(&v_free[i], 1);
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) { mbar_init(&bar_s[t], 1); mbar_init(&bar_p[t], 4); mbar_init(&bar_pv[t], 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) { FF_STAMP(1); FF_STAMP(4); }

  // Running counters, kept alike by every role: it = items done by this CTA (bar_q / q_free phase),
  // kc = K/V blocks streamed so far (ring stage kc % 3, fill kc / 3), c_t = softmax steps of tile t so far
  // (phase of bar_s[t], bar_p[t], bar_pv[t]).
  if (warp == W3_TMA && elect_one_sync()) {
    // ---- TMA producer ------------------------------------------------------------------------
    int it = 0, st = 0;
    uint32_t kc = 0, free_par = 1;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int grp, h, b;
      decode(w, grp, h, b);
      const int nt_act = min(NT, n_tiles - grp * NT), nb_max = nb_of(grp, nt_act - 1);
      if (it > 0) mbar_wait(q_free, (it - 1) & 1);        // the previous item's last S has read the Q tiles
      mbar_arrive_expect_tx(bar_q, nt_act * L::QT);
      for (int t = 0; t < nt_act; ++t) {
        tma_load_4d(smem + L::Q + t * L::QT, &tq, bar_q, 0, h, (grp * NT + t) * TQ, b);
        if (P1) tma_load_4d(smem + L::Q + t * L::QT + QP0, &tq1, bar_q, 64, h, (grp * NT + t) * TQ, b);
      }
#pragma unroll 1
      for (int j = 0; j < nb_max; ++j, ++kc) {
        uint8_t* kd = smem + L::KR + st * L::KST;
        uint8_t* vd = smem + L::VR + st * L::KST;
        if (kc >= NST) mbar_wait(&k_free[st], free_par);
        mbar_arrive_expect_tx(&bar_k[st], L::KST);
        tma_load_4d(kd, &tk, &bar_k[st], 0, h, j * KB, b);
        if (P1) tma_load_4d(kd + KP0, &tk1, &bar_k[st], 64, h, j * KB, b);
        if (kc >= NST) mbar_wait(&v_free[st], free_par);
        mbar_arrive_expect_tx(&bar_v[st], L::KST);
        tma_load_4d(vd, &tv, &bar_v[st], 0, h, j * KB, b);
        if (P1) tma_load_4d(vd + KP0, &tv1, &bar_v[st], 64, h, j * KB, b);
        if (st == NST - 1) { st = 0; free_par ^= 1; } else { ++st; }
      }
    }
  } else if (warp == W3_MMA && elect_one_sync()) {
    // ---- MMA issuer: per key block j, round robin over the tiles: S_t(j), then PV_t(j-1) ------------
    constexpr uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);
    constexpr uint32_t idesc_o0 = make_idesc(TQ, 64, 0, 1);
    constexpr uint32_t idesc_o1 = make_idesc(TQ, 16, 0, 1);
    const uint32_t su = smem_u32(smem);
    auto issue_s = [&](int t, int kst) {            // S_t = Q_t K^T from K stage kst
      const uint32_t k_u = su + L::KR + kst * L::KST, q_u = su + L::Q + t * L::QT;
      const uint32_t d = tmem + S_COL + t * 64;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss(d, make_smem_desc(q_u + k4 * 32, 16, 1024), make_smem_desc(k_u + k4 * 32, 16, 1024), idesc_s, k4 > 0);
      if (P1) umma_ss(d, make_smem_desc32(q_u + QP0, 16, 256), make_smem_desc32(k_u + KP0, 16, 256), idesc_s, 1);
      umma_commit(&bar_s[t]);
    };
    auto issue_pv = [&](int t, int vst, bool acc) {  // O_t (+)= P_t V from V stage vst
      const uint32_t v_u = su + L::VR + vst * L::KST, p_u = su + L::P + t * PB;
      const uint32_t p_t = tmem + P_COL + t * P_STRIDE;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        if (PT)
          umma_ts(tmem + O0_COL + t * 64, p_t + k4 * 8, make_smem_desc(v_u + k4 * 2048, 1024, 1024), idesc_o0,
                  (acc || k4 > 0));
        else
          umma_ss(tmem + O0_COL + t * 64, make_smem_desc(p_u + k4 * 32, 16, 1024),
                  make_smem_desc(v_u + k4 * 2048, 1024, 1024), idesc_o0, (acc || k4 > 0));
      }
      if (P1) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          if (PT)
            umma_ts(tmem + O1_COL + t * 16, p_t + k4 * 8, make_smem_desc32(v_u + KP0 + k4 * 512, 256, 256),
                    idesc_o1, (acc || k4 > 0));
          else
            umma_ss(tmem + O1_COL + t * 16, make_smem_desc(p_u + k4 * 32, 16, 1024),
                    make_smem_desc32(v_u + KP0 + k4 * 512, 256, 256), idesc_o1, (acc || k4 > 0));
        }
      }
      umma_commit(&bar_pv[t]);
    };
    int it = 0, st = 0, vst = NST - 1;             // K stage of block j; V stage of block j - 1
    uint32_t kpar = 0, vpar = 1;                   // fill parities of those
    uint32_t c[NT] = {0u, 0u, 0u};                 // steps started per tile (phase of bar_p[t] for the previous step = c - 1)
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int grp, h, b;
      decode(w, grp, h, b);
      const int nt_act = min(NT, n_tiles - grp * NT);
      const int nb0 = nb_of(grp, 0), nb1 = nb_of(grp, 1), nb2 = nb_of(grp, 2);
      auto nbt = [&](int t) { return t == 0 ? nb0 : (t == 1 ? nb1 : nb2); };
      const int nb_max = nbt(nt_act - 1);
      mbar_wait(bar_q, it & 1);
#pragma unroll 1
      for (int j = 0; j < nb_max; ++j) {
        mbar_wait(&bar_k[st], kpar);
        if (j > 0) mbar_wait(&bar_v[vst], vpar);
        tcgen05_fence_after();
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (t < nt_act && j - 1 < nbt(t)) {                 // tile t still has work at this block
            if (j > 0) {                                      // P_t(j-1) stored, S_t consumed
              mbar_wait(&bar_p[t], (c[t] - 1) & 1);
              tcgen05_fence_after();
            }
            if (P_ALIAS) {                                    // P_t lives in S_t's columns: PV first
              if (j > 0) issue_pv(t, vst, j > 1);
              if (j < nbt(t)) { issue_s(t, st); ++c[t]; }
            } else {
              if (j < nbt(t)) { issue_s(t, st); ++c[t]; }
              if (j > 0) issue_pv(t, vst, j > 1);
            }
            if (it == 0 && j < 4) FF_STAMP(8 + t * 16 + j * 4 + 3);
          }
        }
        umma_commit(&k_free[st]);
        if (j > 0) umma_commit(&v_free[vst]);
        if (j + 1 == nb_max) umma_commit(q_free);             // every S of this item has been issued
        vst = st; vpar = kpar;
        if (st == NST - 1) { st = 0; kpar ^= 1; } else { ++st; }
      }
      // the last PV of the tiles that run to nb_max (shorter tiles issued theirs inside the loop)
      mbar_wait(&bar_v[vst], vpar);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (t < nt_act && nbt(t) == nb_max) {
          mbar_wait(&bar_p[t], (c[t] - 1) & 1);
          tcgen05_fence_after();
          issue_pv(t, vst, nb_max > 1);
        }
      }
      umma_commit(&v_free[vst]);
    }
  } else if (warp < NT * 4) {
    // ---- softmax threads of tile t: one query row per thread -----------------------------------
    const int t = warp >> 2, r = tid - t * TQ;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* sP = smem + L::P + t * PB;
    const float tau = 8.f / a.scale_log2, sl2 = a.scale_log2;
    float f0[32], f1[32];
    uint32_t* s0 = reinterpret_cast<uint32_t*>(f0);
    uint32_t* s1 = reinterpret_cast<uint32_t*>(f1);
    uint32_t c = 0;                                   // steps of this tile so far (all items)
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int grp, h, b;
      decode(w, grp, h, b);
      if (t >= min(NT, n_tiles - grp * NT)) continue;   // this tile does not exist in the group
      const int row0 = (grp * NT + t) * TQ, row = row0 + r;
      const int nb = nb_of(grp, t);
      const bool valid = row < a.Lq;
      const bool active = row0 + ((warp & 3) << 5) < a.Lq;
      float m_ref = -INFINITY, sum = 0.f;
      for (int j = 0; j < nb; ++j, ++c) {
        const uint64_t vis = (active && valid) ? visible<CAUSAL>(a, b, j, row) : 0ull;
        const bool any = __any_sync(0xffffffffu, vis != 0ull);
        mbar_wait(&bar_s[t], c & 1);
        tcgen05_fence_after();
        if (r == 0 && it == 0 && j < 4) FF_STAMP(8 + t * 16 + j * 4 + 0);
        float alpha = 1.f;
        bool grow = false;
        uint4 pk[8];
        if (any) {
          const uint32_t sc = lane_addr + S_COL + t * 64;
          tmem_ld32(sc, s0);
          tmem_ld32(sc + 32, s1);
          tmem_ld_wait();
          if (vis != ~0ull) {
            const uint32_t v0 = (uint32_t)vis, v1 = (uint32_t)(vis >> 32);
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) {
              if (!(v0 & (1u << cc))) f0[cc] = -INFINITY;
              if (!(v1 & (1u << cc))) f1[cc] = -INFINITY;
            }
          }
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) mx[cc & 3] = fmaxf(mx[cc & 3], fmaxf(f0[cc], f1[cc]));
          const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          if (bm > m_ref + tau || (m_ref == -INFINITY && bm > -INFINITY)) {
            grow = m_ref > -INFINITY;
            alpha = grow ? exp2f((m_ref - bm) * sl2) : 1.f;
            m_ref = bm;
          }
          const float nms = m_ref > -INFINITY ? -m_ref * sl2 : 0.f;
          float ps[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
          for (int cc = 0; cc < 32; cc += 2) {
            const float e0 = exp2f(fmaf(f0[cc], sl2, nms)), e1 = exp2f(fmaf(f0[cc + 1], sl2, nms));
            const float g0 = exp2f(fmaf(f1[cc], sl2, nms)), g1 = exp2f(fmaf(f1[cc + 1], sl2, nms));
            ps[(cc >> 1) & 3] += (e0 + e1) + (g0 + g1);
            pw[cc >> 1] = pack2(e0, e1);
            pw[16 + (cc >> 1)] = pack2(g0, g1);
          }
          sum = sum * alpha + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
        } else {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) pk[cc] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (r == 0 && it == 0 && j < 4) FF_STAMP(8 + t * 16 + j * 4 + 1);
        if (j > 0) {                                   // sP_t and O_t are free once PV_t(j-1) has finished
          mbar_wait(&bar_pv[t], (c - 1) & 1);
          tcgen05_fence_after();
          if (__any_sync(0xffffffffu, grow)) {
            uint32_t tt[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              tmem_ld32(lane_addr + O0_COL + t * 64 + half * 32, tt);
              tmem_ld_wait();
#pragma unroll
              for (int cc = 0; cc < 32; ++cc) tt[cc] = __float_as_uint(__uint_as_float(tt[cc]) * alpha);
              tmem_st32(lane_addr + O0_COL + t * 64 + half * 32, tt);
            }
            if (P1) {
              tmem_ld16(lane_addr + O1_COL + t * 16, tt);
              tmem_ld_wait();
#pragma unroll
              for (int cc = 0; cc < 16; ++cc) tt[cc] = __float_as_uint(__uint_as_float(tt[cc]) * alpha);
              tmem_st16(lane_addr + O1_COL + t * 16, tt);
            }
            tmem_st_wait();
          }
        }
        if (PT) {
          if (active) tmem_st32(lane_addr + P_COL + t * P_STRIDE, reinterpret_cast<const uint32_t*>(pk));
          tmem_st_wait();
        } else {
          if (active) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) *reinterpret_cast<uint4*>(sP + sw128_offset(r, cc)) = pk[cc];
          }
          fence_proxy_async_smem();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[t]);
        if (r == 0 && it == 0 && j < 4) FF_STAMP(8 + t * 16 + j * 4 + 2);
      }
      mbar_wait(&bar_pv[t], (c - 1) & 1);
      tcgen05_fence_after();
      if (tid == 0 && it == 0) FF_STAMP(2);
      if (active) {
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        __nv_bfloat16* orow = a.o + (int64_t)b * a.o_bs + (int64_t)row * a.o_rs + h * DH;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(lane_addr + O0_COL + t * 64 + half * 32, s0);
          tmem_ld_wait();
          if (valid) store_bf16(orow + half * 32, s0, 32, inv);
        }
        if (P1) {
          tmem_ld16(lane_addr + O1_COL + t * 16, s0);
          tmem_ld_wait();
          if (valid) store_bf16(orow + 64, s0, 16, inv);
        }
        if (valid) a.lse[((int64_t)b * a.H + h) * a.Lq + row] = sum > 0.f ? m_ref * a.scale + logf(sum) : -INFINITY;
      }
      tcgen05_fence_before();     // orders the O reads above before this warp's next bar_p arrival
      if (tid == 0 && it == 0) FF_STAMP(3);
    }
    if (tid == 0) FF_STAMP(5);
  }
  __syncthreads();
  if (warp == W3_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

static unsigned long long* g_dbg = nullptr;

}  // namespace ff

// ---- host -----------------------------------------------------------------------------------
int make_tmap_tiled_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// q, k, v: (B, L, H, dh) views given by (batch, row, head) strides in elements.
template <int DH, bool CAUSAL>
static int launch_flash(const void* q, int64_t q_bs, int64_t q_rs, int64_t q_hs, const void* k, int64_t k_bs,
                        int64_t k_rs, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                        const uint32_t* kbits, void* o, int64_t o_bs,
                        int64_t o_rs, float* lse, int B, int Lq, int Lk, int H, float scale, cudaStream_t st) {
  using L = ff::Smem<DH>;
  CUtensorMap m[6];
  auto mk = [&](CUtensorMap* out, const void* base, int64_t bs, int64_t rs, int64_t hs, int Ln, int rows,
                bool p1) -> int {
    if (B == 1) bs = rs * (int64_t)Ln;
    const uint64_t dims[4] = {(uint64_t)DH, (uint64_t)H, (uint64_t)Ln, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)hs * 2, (uint64_t)rs * 2, (uint64_t)bs * 2};
    const uint32_t box[4] = {p1 ? 16u : 64u, 1, (uint32_t)rows, 1};
    return make_tmap_tiled_sw(out, base, 4, dims, strides, box, p1 ? 32 : 128);
  };
  int rc;
  if ((rc = mk(&m[0], q, q_bs, q_rs, q_hs, Lq, ff::TQ, false))) return rc;
  if ((rc = mk(&m[1], k, k_bs, k_rs, k_hs, Lk, ff::KB, false))) return rc;
  if ((rc = mk(&m[2], v, v_bs, v_rs, v_hs, Lk, ff::KB, false))) return rc;
  if (L::P1) {
    if ((rc = mk(&m[3], q, q_bs, q_rs, q_hs, Lq, ff::TQ, true))) return rc;
    if ((rc = mk(&m[4], k, k_bs, k_rs, k_hs, Lk, ff::KB, true))) return rc;
    if ((rc = mk(&m[5], v, v_bs, v_rs, v_hs, Lk, ff::KB, true))) return rc;
  } else {
    m[3] = m[0]; m[4] = m[1]; m[5] = m[2];
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ff::flash_fwd_kernel<DH, CAUSAL, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ff::flash_fwd_kernel<DH, CAUSAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)L::TOTAL);
    if (e != cudaSuccess) { set_error("flash_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(ff::flash_fwd_kernel<DH, CAUSAL, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(ff::flash_fwd_kernel<DH, CAUSAL, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  // Which of the two kernels?  The persistent three-tile kernel wins once there are at least two items
  // per SM and more than one query tile (configs[2]: ViT 81 -> 68 us, K4 105 -> 90 us); below that its
  // one-CTA-per-SM quantisation loses to the two-CTA kernel (configs[1]: ViT 22 vs 26 us, K4 20 vs 21 us).
  // UNIMP_FLASH3 = 0 / 1 forces one or the other (A/B runs).
  static const int force3 = getenv("UNIMP_FLASH3") ? atoi(getenv("UNIMP_FLASH3")) : -1;
  // P through tensor memory (default) or through shared memory (UNIMP_FLASH_PT=0, A/B runs)
  static const bool pt = !(getenv("UNIMP_FLASH_PT") && atoi(getenv("UNIMP_FLASH_PT")) == 0);
  static int n_sms = 0;
  if (!n_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sms <= 0)
      n_sms = UNIMP_NUM_SMS;
  }
  const int n_tiles = (Lq + ff::TQ - 1) / ff::TQ;
  const int64_t n_items = (int64_t)((n_tiles + ff::NT - 1) / ff::NT) * H * B;
  const bool three = force3 >= 0 ? force3 != 0 : (n_tiles >= 2 && n_items >= 2 * (int64_t)n_sms);
  using L3 = ff::Smem3<DH>;
  static bool attr3 = false;
  if (three && !attr3) {
    cudaError_t e = cudaFuncSetAttribute(ff::flash_fwd3_kernel<DH, CAUSAL, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L3::TOTAL);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ff::flash_fwd3_kernel<DH, CAUSAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)L3::TOTAL);
    if (e != cudaSuccess) { set_error("flash_fwd3: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr3 = true;
  }
  ff::Args a{};
  a.dbg = ff::g_dbg;
  a.o = (__nv_bfloat16*)o; a.o_bs = o_bs; a.o_rs = o_rs; a.lse = lse;
  a.kbits = kbits; a.kwords = 2 * ((Lk + 63) / 64);
  a.Lq = Lq; a.Lk = Lk; a.H = H; a.Bt = B; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  if (three) {
    const unsigned grid3 = (unsigned)(n_items < n_sms ? n_items : n_sms);    // persistent: one CTA per SM
    if (pt)
      ff::flash_fwd3_kernel<DH, CAUSAL, true><<<grid3, ff::THREADS3, L3::TOTAL, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], a);
    else
      ff::flash_fwd3_kernel<DH, CAUSAL, false><<<grid3, ff::THREADS3, L3::TOTAL, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], a);
    UNIMP_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid(n_tiles, H, B);
  if (pt)
    ff::flash_fwd_kernel<DH, CAUSAL, true><<<grid, ff::THREADS, L::TOTAL, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], a);
  else
    ff::flash_fwd_kernel<DH, CAUSAL, false><<<grid, ff::THREADS, L::TOTAL, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], a);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

// dh = 64, unmasked (Perceiver / ViT): q (B,Lq,H,64), k/v (B,Lk,H,64) with head stride 64
int launch_flash_fwd_64(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_mview_t o, float* lse, int B,
                        int Lq, int Lk, int H, float scale, cudaStream_t st) {
  return launch_flash<64, false>(q.ptr, q.batch_stride, q.row_stride, 64, k.ptr, k.batch_stride, k.row_stride, 64,
                                 v.ptr, v.batch_stride, v.row_stride, 64, nullptr, o.ptr, o.batch_stride,
                                 o.row_stride, lse, B, Lq, Lk, H, scale, st);
}

// dh = 80, causal + key bits (GPT-NeoX)
int launch_flash_fwd_80(const void* q, const void* k, const void* v, int64_t bs, int64_t rs, int64_t hs,
                        const uint32_t* kbits, void* o, float* lse, int B, int T, int H, float scale,
                        cudaStream_t st) {
  return launch_flash<80, true>(q, bs, rs, hs, k, bs, rs, hs, v, bs, rs, hs, kbits, o, (int64_t)T * H * 80,
                                (int64_t)H * 80, lse, B, T, T, H, scale, st);
}

}  // namespace unimp

// Test hook (not in the public header): device buffer of 64 x u64 per CTA that the next flash_fwd
// launches fill with %globaltimer stamps; NULL switches it off.
extern "C" void unimp__flash_fwd_debug(unsigned long long* buf) { unimp::ff::g_dbg = buf; }

// Test hook: resident CTAs per SM the runtime grants the kernel (dh = 64 or 80).
extern "C" int unimp__flash_fwd_occupancy(int dh) {
  int n = -1;
  if (dh == 80) {
    cudaFuncSetAttribute(unimp::ff::flash_fwd_kernel<80, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)unimp::ff::Smem<80>::TOTAL);
    cudaFuncSetAttribute(unimp::ff::flash_fwd_kernel<80, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, unimp::ff::flash_fwd_kernel<80, true, false>, unimp::ff::THREADS,
                                                  unimp::ff::Smem<80>::TOTAL);
  } else {
    cudaFuncSetAttribute(unimp::ff::flash_fwd_kernel<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)unimp::ff::Smem<64>::TOTAL);
    cudaFuncSetAttribute(unimp::ff::flash_fwd_kernel<64, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, unimp::ff::flash_fwd_kernel<64, false, false>, unimp::ff::THREADS,
                                                  unimp::ff::Smem<64>::TOTAL);
  }
  return n;
}
