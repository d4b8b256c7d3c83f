// K4: causal self-attention of the (frozen) GPT-NeoX decoder layers, bf16, head dim 80, on tcgen05.
//
// Replaces the attention core of HF `GPTNeoXAttention.forward` (transformers gpt_neox; reached from
// upstream `FlamingoLayer.forward` -> `decoder_layer(...)`, call site UniMP/mmrec.py:177-181): after
// rotary, softmax(scale * q k^T + causal/padding mask) v per head.  SURVEY.md §8 row f3.
//
// Head dim 80 on SWIZZLE_128B operands: a head's 80 columns are read as TWO 64-column panels from a
// tensor map whose innermost extent is 80 — panel 1 (columns 64..127) gets columns 80..127 zero
// filled by the TMA unit, no padded copy of q/k/v exists in memory.  S = Q K^T takes 4 K-steps from
// panel 0 and ONE from panel 1 (K = 80 exactly); O and the gradients are produced per panel
// (N = 64, of which 16 columns are used in panel 1).
//
// Mask: key j is visible to query i iff j <= i and key_bits[b][j] (bit j%32 of word j/32; NULL = all
// keys valid) — what HF builds from a 2-D attention_mask for a causal LM.  Rows that see no key
// get o = 0 / lse = -inf.
//
// Forward : CTA = (128-query tile, head, sample), two sweeps over the tile's key blocks like
//           attn_fwd2_tc_kernel (row max, then exp + PV with S recomputed); 2 CTAs per SM.
// Backward: CTA = (64-key block, head, sample) walks the query tiles at or below the diagonal;
//           dK/dV accumulate in TMEM across tiles; each pair's dQ tile is staged in shared memory and
//           added into an fp32 buffer (B,H,Tp,84) with ONE bulk reduce (cp.reduce.async.bulk .add.f32,
//           43 KB, asynchronous) — per-thread vector atomics were L2-bound (917 us at T=1024).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace unimp {
namespace lm {

using namespace tc;

constexpr int TQ = 128, KB = 64, DH = 80;
constexpr uint32_t QP = TQ * 128;     // one [128 rows][64 cols] bf16 panel, 16 KB
constexpr uint32_t KP = KB * 128;     // one [64 rows][64 cols] panel, 8 KB
constexpr uint32_t PB = TQ * 128;     // P / dS: [128 rows][64 keys], 16 KB
constexpr int THREADS = TQ + 32;

struct Args {
  __nv_bfloat16* o;          // fwd: out (B,T,H*80); bwd: forward output (read)
  const __nv_bfloat16* d_o;  // bwd
  int64_t o_bs, o_rs;        // shared by o and d_o (both (B,T,H*80) contiguous views)
  float* lse;                // (B,H,T)
  const uint32_t* kbits;     // (B, kwords) or NULL
  int kwords;
  float* dq32;               // bwd: (B,H,Tp,84) fp32 (Tp = T rounded up to 128), zeroed by the launcher
  __nv_bfloat16 *dk, *dv;    // bwd: (B,T,H,80) bf16 contiguous
  int T, H;
  float scale, scale_log2;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// 32 floats -> half of a 128-byte swizzled row of a [128][64] bf16 tile
__device__ __forceinline__ void store_half(uint8_t* tile, int row, int half, const float* p) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    v.x = pack2(p[8 * c + 0], p[8 * c + 1]);
    v.y = pack2(p[8 * c + 2], p[8 * c + 3]);
    v.z = pack2(p[8 * c + 4], p[8 * c + 5]);
    v.w = pack2(p[8 * c + 6], p[8 * c + 7]);
    *reinterpret_cast<uint4*>(tile + sw128_offset(row, half * 4 + c)) = v;
  }
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

// Dynamic shared memory only (no static __shared__): the base of the window is 1024-byte aligned,
// which SWIZZLE_128B tiles need, without paying 1 KB of slack (two CTAs per SM fit to the byte).
constexpr uint32_t F_Q0 = 0, F_Q1 = QP, F_RING = 2 * QP, F_STAGE = 4 * KP, F_P = F_RING + 2 * F_STAGE,
                   F_BAR = F_P + PB, F_SMEM = F_BAR + 128;

__global__ void __launch_bounds__(THREADS, 2)
lm_attn_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const Args a) {
  constexpr uint32_t S_COL = 0, O0_COL = 64, O1_COL = 128, TMEM_COLS = 256;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_BAR);
  uint64_t *bar_q = bars, *bar_kv = bars + 1, *bar_s = bars + 3, *bar_p = bars + 4, *bar_pv = bars + 5,
           *bar_o = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = tid < TQ;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;   // long (late) tiles first
  const int row0 = qt * TQ, h = blockIdx.y, b = blockIdx.z;
  const int nb = min((a.T + KB - 1) / KB, (row0 + TQ) / KB), nsteps = 2 * nb;

  // step s: key block s % nb; steps [0, nb) = sweep 1 (K only), [nb, 2nb) = sweep 2 (K and V)
  auto load_step = [&](int s) {
    const int st = s & 1, j = s % nb;
    uint8_t* base = smem + F_RING + st * F_STAGE;
    if (s < nb) {
      mbar_arrive_expect_tx(&bar_kv[st], 2 * KP);
      tma_load_4d(base, &tk, &bar_kv[st], 0, h, j * KB, b);
      tma_load_4d(base + KP, &tk, &bar_kv[st], 64, h, j * KB, b);
    } else {
      mbar_arrive_expect_tx(&bar_kv[st], 4 * KP);
      tma_load_4d(base, &tk, &bar_kv[st], 0, h, j * KB, b);
      tma_load_4d(base + KP, &tk, &bar_kv[st], 64, h, j * KB, b);
      tma_load_4d(base + 2 * KP, &tv, &bar_kv[st], 0, h, j * KB, b);
      tma_load_4d(base + 3 * KP, &tv, &bar_kv[st], 64, h, j * KB, b);
    }
  };

  if (warp == 4) {
    if (elect_one_sync()) {
      if (smem_u32(smem) & 1023u) {
        printf("unimp: lm_attn_fwd: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
      }
      mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_pv, 1);
      mbar_init(bar_o, 1); mbar_init(&bar_kv[0], 1); mbar_init(&bar_kv[1], 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(bar_q, 2 * QP);
      tma_load_4d(smem + F_Q0, &tq, bar_q, 0, h, row0, b);
      tma_load_4d(smem + F_Q1, &tq, bar_q, 64, h, row0, b);
      load_step(0);
      if (nsteps > 1) load_step(1);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);
  const uint32_t idesc_o = make_idesc(TQ, 64, 0, 1);

  if (warp == 4 && elect_one_sync()) {
    const uint32_t q0 = smem_u32(smem + F_Q0), q1 = smem_u32(smem + F_Q1), ring = smem_u32(smem + F_RING),
                   p_u = smem_u32(smem + F_P);
    auto issue_s = [&](uint32_t stage_u) {        // S = Q K^T over K = 80
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss(tmem + S_COL, make_smem_desc(q0 + k4 * 32, 16, 1024), make_smem_desc(stage_u + k4 * 32, 16, 1024),
                idesc_s, k4 > 0);
      umma_ss(tmem + S_COL, make_smem_desc(q1, 16, 1024), make_smem_desc(stage_u + KP, 16, 1024), idesc_s, 1);
      umma_commit(bar_s);
    };
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_kv[0], 0);
    tcgen05_fence_after();
    issue_s(ring);
    uint32_t ring_ph = 0;          // parity of stage 0's current fill; stage 1 lags by one step
    for (int s0 = 0; s0 < nsteps; s0 += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int s = s0 + u;      // stage = u, bar_p parity = u
        if (s < nsteps) {
          const bool sweep2 = s >= nb;
          mbar_wait(bar_p, u);     // sweep 1: S_s consumed; sweep 2: P_s in shared memory
          if (sweep2) {
            tcgen05_fence_after();
            const uint32_t v_u = ring + u * F_STAGE + 2 * KP;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_ss(tmem + O0_COL, make_smem_desc(p_u + k4 * 32, 16, 1024),
                      make_smem_desc(v_u + k4 * 2048, 1024, 1024), idesc_o, (s > nb || k4 > 0));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_ss(tmem + O1_COL, make_smem_desc(p_u + k4 * 32, 16, 1024),
                      make_smem_desc(v_u + KP + k4 * 2048, 1024, 1024), idesc_o, (s > nb || k4 > 0));
            umma_commit(bar_pv);
            if (s + 1 == nsteps) umma_commit(bar_o);
          }
          if (s + 1 < nsteps) {
            mbar_wait(&bar_kv[1 - u], u == 0 ? ring_ph : (ring_ph ^ 1));
            tcgen05_fence_after();
            issue_s(ring + (1 - u) * F_STAGE);
          }
          if (s + 2 < nsteps) {
            if (sweep2) mbar_wait(bar_pv, (s - nb) & 1);   // PV_s has released stage u
            load_step(s + 2);
          }
        }
      }
      ring_ph ^= 1;
    }
  }

  if (worker) {
    const int row = row0 + tid;
    const bool valid = row < a.T;
    const bool active = row0 + (warp << 5) < a.T;        // warp-uniform
    float m = -INFINITY, sum = 0.f, ms = 0.f;
    uint32_t r[32];
    for (int s = 0; s < nsteps; ++s) {
      const int j = s % nb;
      const uint64_t vis = (active && valid) ? visible(a, b, j, row) : 0ull;
      const bool any = __any_sync(0xffffffffu, vis != 0ull);      // warp-uniform
      mbar_wait(bar_s, s & 1);
      tcgen05_fence_after();
      if (s < nb) {
        if (any) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t vh = (uint32_t)(vis >> (32 * half));
            tmem_ld32(lane_addr + S_COL + half * 32, r);
            tmem_ld_wait();
            if (vh == 0xffffffffu) {
#pragma unroll
              for (int c = 0; c < 32; ++c) m = fmaxf(m, __uint_as_float(r[c]));
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if ((vh >> c) & 1u) m = fmaxf(m, __uint_as_float(r[c]));
            }
          }
        }
        if (s + 1 == nb) ms = (m > -INFINITY) ? m * a.scale_log2 : 0.f;
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
      } else {
        if (active) {
          float psum = 0.f;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float p[32];
            const uint32_t vh = (uint32_t)(vis >> (32 * half));
            if (any) {
              tmem_ld32(lane_addr + S_COL + half * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                p[c] = ((vh >> c) & 1u) ? exp2f(__uint_as_float(r[c]) * a.scale_log2 - ms) : 0.f;
                psum += p[c];
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) p[c] = 0.f;
            }
            store_half(smem + F_P, tid, half, p);
          }
          sum += psum;
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
      }
    }
    mbar_wait(bar_o, 0);
    tcgen05_fence_after();
    if (active) {
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      __nv_bfloat16* orow = a.o + (int64_t)b * a.o_bs + (int64_t)row * a.o_rs + h * DH;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld32(lane_addr + O0_COL + half * 32, r);
        tmem_ld_wait();
        if (valid) store_bf16(orow + half * 32, r, 32, inv);
      }
      tmem_ld32(lane_addr + O1_COL, r);
      tmem_ld_wait();
      if (valid) {
        store_bf16(orow + 64, r, 16, inv);
        a.lse[((int64_t)b * a.H + h) * a.T + row] = sum > 0.f ? m * a.scale + logf(sum) : -INFINITY;
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Backward.  Per (query tile i, key block j) pair, with S, dP, P, dS as in attn_bwd_tc_kernel:
//   [dV_j | dK_j] += [P | dS]^T [dO_i | Q_i]   per 64-column panel (two M=N=128 MMAs chains)
//   dQ_i += dS K_j                              per panel; staged as fp32 [128][84] (336-byte pitch:
//                                               conflict-free float4 stores) and bulk-reduced into dq32
// Shared memory: [dO p0][Q p0][dO p1][Q p1] (B operand chunks 16 KB apart), [P][dS], K p0/p1, V p0/p1,
// dQ staging.
// TMEM: S 64 | dP 64 | dQ p0 64 | dQ p1 64 | dKV p0 128 | dKV p1 128 = 512 columns, one CTA per SM.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t B_DO0 = 0, B_Q0 = QP, B_DO1 = 2 * QP, B_Q1 = 3 * QP, B_P = 4 * QP, B_DS = B_P + PB,
                   B_K0 = B_DS + PB, B_K1 = B_K0 + KP, B_V0 = B_K1 + KP, B_V1 = B_V0 + KP,
                   B_STG = B_V1 + KP, DQ_PITCH = 84, STG_BYTES = TQ * DQ_PITCH * 4,
                   B_BAR = B_STG + STG_BYTES, B_SMEM = B_BAR + 128;

__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               :
               : "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(THREADS, 1)
lm_attn_bwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tdo,
                   const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                   const Args a) {
  constexpr uint32_t S_COL = 0, DP_COL = 64, DQ0_COL = 128, DQ1_COL = 192, DKV0_COL = 256, DKV1_COL = 384,
                     TMEM_COLS = 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_BAR);
  uint64_t *bar_qdo = bars, *bar_kv = bars + 1, *bar_s = bars + 2, *bar_g = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool worker = tid < TQ;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nq = (a.T + TQ - 1) / TQ;
  const int q_lo = (kb * KB) / TQ;            // first query tile with a row at or after key kb*64

  if (warp == 4) {
    if (elect_one_sync()) {
      if (smem_u32(smem) & 1023u) {
        printf("unimp: lm_attn_bwd: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
      }
      mbar_init(bar_qdo, 1); mbar_init(bar_kv, 1); mbar_init(bar_s, 1);
      mbar_init(bar_g, 2);     // MMAs of the pair done (tcgen05.commit) + dQ staging free (issuer)
      fence_barrier_init();
      mbar_arrive_expect_tx(bar_kv, 4 * KP);
      tma_load_4d(smem + B_K0, &tk, bar_kv, 0, h, kb * KB, b);
      tma_load_4d(smem + B_K1, &tk, bar_kv, 64, h, kb * KB, b);
      tma_load_4d(smem + B_V0, &tv, bar_kv, 0, h, kb * KB, b);
      tma_load_4d(smem + B_V1, &tv, bar_kv, 64, h, kb * KB, b);
      mbar_arrive_expect_tx(bar_qdo, 4 * QP);
      tma_load_4d(smem + B_DO0, &tdo, bar_qdo, 0, h, q_lo * TQ, b);
      tma_load_4d(smem + B_DO1, &tdo, bar_qdo, 64, h, q_lo * TQ, b);
      tma_load_4d(smem + B_Q0, &tq, bar_qdo, 0, h, q_lo * TQ, b);
      tma_load_4d(smem + B_Q1, &tq, bar_qdo, 64, h, q_lo * TQ, b);
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
  const uint32_t idesc_s = make_idesc(TQ, KB, 0, 0);       // S, dP
  const uint32_t idesc_dq = make_idesc(TQ, 64, 0, 1);      // dQ = dS K   (B = K panel, MN-major)
  const uint32_t idesc_dkv = make_idesc(128, 128, 1, 1);   // [P|dS]^T [dO|Q]
  const float l2e = 1.4426950408889634f;

  uint32_t ph = 0;
  uint32_t r[32];
  float* stg = reinterpret_cast<float*>(smem + B_STG);
  const int Tp = nq * TQ;
  float* dq_head = a.dq32 + ((int64_t)b * a.H + h) * Tp * DQ_PITCH;
  if (worker) *reinterpret_cast<float4*>(stg + tid * DQ_PITCH + DH) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int qt = q_lo; qt < nq; ++qt) {
    const int it = qt - q_lo;
    const int row = qt * TQ + tid;
    const bool valid = worker && row < a.T;
    // ---- S and dP ------------------------------------------------------------------------
    if (warp == 4 && elect_one_sync()) {
      if (it > 0) {            // the previous pair's dQ tile (staged before the loop-end barrier)
        bulk_reduce_add_f32(dq_head + (int64_t)(qt - 1) * TQ * DQ_PITCH, stg, STG_BYTES);
        tma_store_commit();
      }
      if (it == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_qdo, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss(tmem + S_COL, make_smem_desc(su + B_Q0 + k4 * 32, 16, 1024),
                make_smem_desc(su + B_K0 + k4 * 32, 16, 1024), idesc_s, k4 > 0);
      umma_ss(tmem + S_COL, make_smem_desc(su + B_Q1, 16, 1024), make_smem_desc(su + B_K1, 16, 1024), idesc_s, 1);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ss(tmem + DP_COL, make_smem_desc(su + B_DO0 + k4 * 32, 16, 1024),
                make_smem_desc(su + B_V0 + k4 * 32, 16, 1024), idesc_s, k4 > 0);
      umma_ss(tmem + DP_COL, make_smem_desc(su + B_DO1, 16, 1024), make_smem_desc(su + B_V1, 16, 1024), idesc_s, 1);
      umma_commit(bar_s);
    }
    if (worker) {
      // delta = rowsum(dO o O) and lse for this row while the MMAs run
      float delta = 0.f, lse_l2 = 0.f;
      uint64_t vis = 0ull;
      if (valid) {
        const __nv_bfloat16* op = a.o + (int64_t)b * a.o_bs + (int64_t)row * a.o_rs + h * DH;
        const __nv_bfloat16* gp = a.d_o + (int64_t)b * a.o_bs + (int64_t)row * a.o_rs + h * DH;
#pragma unroll
        for (int c = 0; c < DH / 8; ++c) {
          Vec16<__nv_bfloat16> ov, gv;
          float of[8], gf[8];
          ov.load(op + c * 8); gv.load(gp + c * 8);
          ov.unpack(of); gv.unpack(gf);
#pragma unroll
          for (int e = 0; e < 8; ++e) delta = fmaf(of[e], gf[e], delta);
        }
        const float lse = a.lse[((int64_t)b * a.H + h) * a.T + row];
        lse_l2 = lse * l2e;
        if (lse > -INFINITY) vis = visible(a, b, kb, row);
      }
      mbar_wait(bar_s, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t rp[32];
        tmem_ld32(lane_addr + S_COL + half * 32, r);
        tmem_ld32(lane_addr + DP_COL + half * 32, rp);
        tmem_ld_wait();
        const uint32_t vh = (uint32_t)(vis >> (32 * half));
        float pv[32], dsv[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float p = 0.f, ds = 0.f;
          if ((vh >> c) & 1u) {
            p = exp2f(__uint_as_float(r[c]) * a.scale_log2 - lse_l2);
            ds = p * (__uint_as_float(rp[c]) - delta) * a.scale;
          }
          pv[c] = p;
          dsv[c] = ds;
        }
        store_half(smem + B_P, tid, half, pv);
        store_half(smem + B_DS, tid, half, dsv);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
    }
    __syncthreads();
    // ---- [dV|dK] += [P|dS]^T [dO|Q] per panel, dQ = dS K per panel -------------------------
    if (warp == 4 && elect_one_sync()) {
      tcgen05_fence_after();
#pragma unroll
      for (int k8 = 0; k8 < TQ / 16; ++k8)
        umma_ss(tmem + DKV0_COL, make_smem_desc(su + B_P + k8 * 2048, PB, 1024),
                make_smem_desc(su + B_DO0 + k8 * 2048, QP, 1024), idesc_dkv, (it > 0 || k8 > 0));
#pragma unroll
      for (int k8 = 0; k8 < TQ / 16; ++k8)
        umma_ss(tmem + DKV1_COL, make_smem_desc(su + B_P + k8 * 2048, PB, 1024),
                make_smem_desc(su + B_DO1 + k8 * 2048, QP, 1024), idesc_dkv, (it > 0 || k8 > 0));
#pragma unroll
      for (int k4 = 0; k4 < KB / 16; ++k4)
        umma_ss(tmem + DQ0_COL, make_smem_desc(su + B_DS + k4 * 32, 16, 1024),
                make_smem_desc(su + B_K0 + k4 * 2048, 1024, 1024), idesc_dq, k4 > 0);
#pragma unroll
      for (int k4 = 0; k4 < KB / 16; ++k4)
        umma_ss(tmem + DQ1_COL, make_smem_desc(su + B_DS + k4 * 32, 16, 1024),
                make_smem_desc(su + B_K1 + k4 * 2048, 1024, 1024), idesc_dq, k4 > 0);
      umma_commit(bar_g);
      tma_store_wait_read();             // the previous dQ reduce has left the staging buffer
      mbar_arrive(bar_g);
      mbar_wait(bar_g, ph);              // the Q/dO panels are free: fetch the next tile's
      if (qt + 1 < nq) {
        mbar_arrive_expect_tx(bar_qdo, 4 * QP);
        tma_load_4d(smem + B_DO0, &tdo, bar_qdo, 0, h, (qt + 1) * TQ, b);
        tma_load_4d(smem + B_DO1, &tdo, bar_qdo, 64, h, (qt + 1) * TQ, b);
        tma_load_4d(smem + B_Q0, &tq, bar_qdo, 0, h, (qt + 1) * TQ, b);
        tma_load_4d(smem + B_Q1, &tq, bar_qdo, 64, h, (qt + 1) * TQ, b);
      }
    }
    if (worker) {
      mbar_wait(bar_g, ph);
      tcgen05_fence_after();
      float* dst = stg + tid * DQ_PITCH;
#pragma unroll
      for (int part = 0; part < 3; ++part) {      // dQ p0 cols 0-31, 32-63, p1 cols 0-15
        tmem_ld32(lane_addr + (part < 2 ? DQ0_COL + part * 32 : DQ1_COL), r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < (part < 2 ? 8 : 4); ++c)
          *reinterpret_cast<float4*>(dst + part * 32 + c * 4) =
              make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                          __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
    }
    ph ^= 1;
    __syncthreads();   // TMEM reads done before the next pair's MMAs overwrite S / dP / dQ
    tcgen05_fence_after();
  }

  if (warp == 4 && elect_one_sync()) {          // the last pair's dQ tile
    bulk_reduce_add_f32(dq_head + (int64_t)(nq - 1) * TQ * DQ_PITCH, stg, STG_BYTES);
    tma_store_commit();
    bulk_wait_all();
  }
  // ---- flush dV (lanes 0-63) and dK (lanes 64-127) of this key block -----------------------
  if (worker) {
    const int key = kb * KB + (tid & 63);
    const bool is_k = tid >= 64;
    __nv_bfloat16* dst = (is_k ? a.dk : a.dv) + (((int64_t)b * a.T + key) * a.H + h) * DH;
    const uint32_t c0 = DKV0_COL + (is_k ? 64 : 0), c1 = DKV1_COL + (is_k ? 64 : 0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tmem_ld32(lane_addr + c0 + half * 32, r);
      tmem_ld_wait();
      if (key < a.T) store_bf16(dst + half * 32, r, 32, 1.f);
    }
    tmem_ld32(lane_addr + c1, r);
    tmem_ld_wait();
    if (key < a.T) store_bf16(dst + 64, r, 16, 1.f);
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
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
static int make_map(CUtensorMap* out, const void* base, int64_t bs, int64_t rs, int64_t hs, int B, int T,
                    int H, int box_rows) {
  if (B == 1) bs = rs * (int64_t)T;
  const uint64_t dims[4] = {(uint64_t)DH, (uint64_t)H, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)hs * 2, (uint64_t)rs * 2, (uint64_t)bs * 2};
  const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_tiled(out, base, 4, dims, strides, box);
}

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
  // one-sweep kernel (flash_fwd.cu); UNIMP_LM_FWD1=1 keeps the two-sweep kernel below for A/B runs
  static const bool two_sweep = getenv("UNIMP_LM_FWD1") && atoi(getenv("UNIMP_LM_FWD1")) != 0;
  if (!two_sweep)
    return launch_flash_fwd_80(q, k, v, batch_stride, row_stride, head_stride, key_bits, o, lse, B, T, H, scale,
                               (cudaStream_t)stream);
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = lm::make_map(&tq, q, batch_stride, row_stride, head_stride, B, T, H, lm::TQ))) return rc;
  if ((rc = lm::make_map(&tk, k, batch_stride, row_stride, head_stride, B, T, H, lm::KB))) return rc;
  if ((rc = lm::make_map(&tv, v, batch_stride, row_stride, head_stride, B, T, H, lm::KB))) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lm::lm_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)lm::F_SMEM);
    if (e != cudaSuccess) { set_error("lm_attn_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(lm::lm_attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  lm::Args a{};
  a.o = (__nv_bfloat16*)o; a.o_bs = (int64_t)T * H * dh; a.o_rs = (int64_t)H * dh;
  a.lse = lse; a.kbits = key_bits; a.kwords = 2 * ((T + 63) / 64);
  a.T = T; a.H = H; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((T + lm::TQ - 1) / lm::TQ, H, B);
  lm::lm_attn_fwd_kernel<<<grid, lm::THREADS, lm::F_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, a);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_lm_attn_bwd(const void* q, const void* k, const void* v, int64_t batch_stride,
                                 int64_t row_stride, int64_t head_stride, const uint32_t* key_bits,
                                 const void* o, const void* d_o, const float* lse, float* dq32, void* dk,
                                 void* dv, int B, int T, int H, int dh, float scale, int dtype,
                                 void* stream) {
  UNIMP_CHECK_ARG(q && k && v && o && d_o && lse && dq32 && dk && dv, UNIMP_E_NULL,
                  "unimp_lm_attn_bwd: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0 && H > 0, UNIMP_E_SHAPE, "unimp_lm_attn_bwd: empty shape");
  const char* why = lm::unsupported(q, k, v, batch_stride, row_stride, head_stride, dh, dtype);
  UNIMP_CHECK_ARG(!why, UNIMP_E_SHAPE, "unimp_lm_attn_bwd: %s", why);
  UNIMP_CHECK_ARG(aligned16(o) && aligned16(d_o) && aligned16(dq32) && aligned16(dk) && aligned16(dv),
                  UNIMP_E_ALIGN, "unimp_lm_attn_bwd: o/d_o/dq32/dk/dv must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tq, tdo, tk, tv;
  int rc;
  const int64_t D = (int64_t)H * dh;
  if ((rc = lm::make_map(&tq, q, batch_stride, row_stride, head_stride, B, T, H, lm::TQ))) return rc;
  if ((rc = lm::make_map(&tdo, d_o, (int64_t)T * D, D, dh, B, T, H, lm::TQ))) return rc;
  if ((rc = lm::make_map(&tk, k, batch_stride, row_stride, head_stride, B, T, H, lm::KB))) return rc;
  if ((rc = lm::make_map(&tv, v, batch_stride, row_stride, head_stride, B, T, H, lm::KB))) return rc;
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
  lm::Args a{};
  a.o = (__nv_bfloat16*)const_cast<void*>(o); a.d_o = (const __nv_bfloat16*)d_o;
  a.o_bs = (int64_t)T * D; a.o_rs = D;
  a.lse = const_cast<float*>(lse); a.kbits = key_bits; a.kwords = 2 * ((T + 63) / 64);
  a.dq32 = dq32; a.dk = (__nv_bfloat16*)dk; a.dv = (__nv_bfloat16*)dv;
  a.T = T; a.H = H; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((T + lm::KB - 1) / lm::KB, H, B);
  lm::lm_attn_bwd_kernel<<<grid, lm::THREADS, lm::B_SMEM, st>>>(tq, tdo, tk, tv, a);
  UNIMP_CHECK_LAUNCH();
  return 0;
}
