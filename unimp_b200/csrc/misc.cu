// Integer/bookkeeping kernels on either side of the attention path, plus the fused AdamW:
//   unimp_text_time    (a4)  media_locations -> per-token image index, ballot prefix scan
//   unimp_mask_labels  (a9)  reference UniMP/mmrec.py:143-168 as a warp scan
//   unimp_adamw_step   (f1)  reference UniMP/mmrec.py:671 (+ clip 247-248), one HBM pass
//   unimp_sumsq        grad-norm partial for the clip
#include "common.cuh"

namespace unimp {

// One warp per sample; 32 tokens per step; running count carried in a register.
__global__ void text_time_kernel(const int64_t* __restrict__ lang_x, int64_t media_id, int B, int T,
                                 int use_cached, int T_out, int32_t* __restrict__ text_time) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int64_t* row = lang_x + (int64_t)b * T;
  int run = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool f = t < T && row[t] == media_id;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (!use_cached && t < T)
      text_time[(int64_t)b * T_out + t] = run + __popc(m & (0xffffffffu >> (31 - lane)));
    run += __popc(m);
  }
  if (use_cached)
    for (int t = lane; t < T_out; t += 32) text_time[(int64_t)b * T_out + t] = run;
}

// State after a token depends only on the LAST event token (<answer> sets, <|endofchunk|>
// clears), so "in answer span before token j" == last <answer> before j is later than the
// last <|endofchunk|> before j.  One warp per sample.
__global__ void mask_labels_kernel(const int64_t* __restrict__ ids, int64_t answer_id,
                                   int64_t eoc_id, int64_t media_id, int64_t pad_id,
                                   int64_t* __restrict__ labels, int B, int T) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int64_t* row = ids + (int64_t)b * T;
  int last_ans = -1, last_eoc = -1;  // positions strictly before the current chunk
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const int64_t tok = t < T ? row[t] : pad_id;
    const unsigned ma = __ballot_sync(0xffffffffu, t < T && tok == answer_id);
    const unsigned me = __ballot_sync(0xffffffffu, t < T && tok == eoc_id);
    const unsigned below = lane ? (0xffffffffu >> (32 - lane)) : 0u;
    const unsigned pa = ma & below, pe = me & below;
    const int la = pa ? t0 + 31 - __clz(pa) : last_ans;
    const int le = pe ? t0 + 31 - __clz(pe) : last_eoc;
    const bool in_answer = la > le;
    if (t < T) {
      int64_t lab = tok;
      if (!in_answer || tok == eoc_id) lab = -100;     // mmrec.py:149-156
      if (tok == pad_id) lab = -100;                   // mmrec.py:157
      if (t == 0) lab = -100;                          // mmrec.py:158
      if (tok == answer_id || tok == media_id) lab = -100;  // mmrec.py:167-168
      labels[(int64_t)b * T + t] = lab;
    }
    if (ma) last_ans = t0 + 31 - __clz(ma);
    if (me) last_eoc = t0 + 31 - __clz(me);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
sumsq_kernel(const T* __restrict__ g, int64_t n, float* __restrict__ acc) {
  constexpr int N = Vec16<T>::N;
  const int64_t nvec = n / N;
  float s = 0.f;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nvec;
       j += (int64_t)gridDim.x * blockDim.x) {
    Vec16<T> v;
    float f[N];
    v.load_stream(g + j * N);
    v.unpack(f);
#pragma unroll
    for (int i = 0; i < N; ++i) s += f[i] * f[i];
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * N + threadIdx.x; i < n; i += blockDim.x) {
      const float f = Elem<T>::to_f(g[i]);
      s += f * f;
    }
  __shared__ float sh[32];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

// One pass: read grad, master, m, v; write master, m, v and the T working copy.
// U = vectors (4 parameters each) a thread keeps in flight per iteration: 1 for the resident
// grid-stride wave; 4 for the BACKGROUND geometry (2 small CTAs per SM that leave room for
// concurrently running compute-bound kernels: every load of the 4 vectors is issued before the
// first use, 4 x 56 B per thread in flight).
template <typename T>
__device__ __forceinline__ void adamw_update4(float* p, float* mm, float* vv, const float* g4, float clip,
                                              float lr, float wd, float b1, float b2, float eps, float bc1,
                                              float bc2_sqrt) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float g = g4[i] * clip;
    p[i] *= 1.f - lr * wd;                       // decoupled weight decay
    mm[i] = b1 * mm[i] + (1.f - b1) * g;
    vv[i] = b2 * vv[i] + (1.f - b2) * g * g;
    const float denom = sqrtf(vv[i]) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mm[i] / denom);
  }
}

template <typename T, int U>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ master, T* __restrict__ param, const T* __restrict__ grad,
             float* __restrict__ m, float* __restrict__ v, int64_t n, const float* __restrict__ hyper,
             float b1, float b2, float eps, float wd, const float* __restrict__ gnorm_sq,
             float max_norm, float grad_scale) {
  const float lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];
  float clip = grad_scale;
  if (gnorm_sq) {
    // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max=1)
    const float norm = sqrtf(*gnorm_sq) * grad_scale;
    clip *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += stride * U) {
    float4 p4[U], m4[U], v4[U];
    float g4[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i + 4 <= n) {
        p4[u] = *reinterpret_cast<float4*>(master + i);
        m4[u] = *reinterpret_cast<float4*>(m + i);
        v4[u] = *reinterpret_cast<float4*>(v + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) g4[u][e] = Elem<T>::to_f(grad[i + e]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i + 4 <= n) {
        float p[4] = {p4[u].x, p4[u].y, p4[u].z, p4[u].w}, mm[4] = {m4[u].x, m4[u].y, m4[u].z, m4[u].w},
              vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
        adamw_update4<T>(p, mm, vv, g4[u], clip, lr, wd, b1, b2, eps, bc1, bc2_sqrt);
#pragma unroll
        for (int e = 0; e < 4; ++e) param[i + e] = Elem<T>::from_f(p[e]);
        *reinterpret_cast<float4*>(master + i) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      } else if (i < n) {
        for (int64_t j = i; j < n; ++j) {
          const float g = Elem<T>::to_f(grad[j]) * clip;
          float p = master[j] * (1.f - lr * wd);
          const float mm = b1 * m[j] + (1.f - b1) * g;
          const float vv = b2 * v[j] + (1.f - b2) * g * g;
          p -= (lr / bc1) * (mm / (sqrtf(vv) / bc2_sqrt + eps));
          master[j] = p; m[j] = mm; v[j] = vv;
          param[j] = Elem<T>::from_f(p);
        }
      }
    }
  }
}

}  // namespace unimp

using namespace unimp;

extern "C" int unimp_text_time(const int64_t* lang_x, int64_t media_token_id, int B, int T,
                               int use_cached, int T_out, int32_t* text_time, void* stream) {
  UNIMP_CHECK_ARG(lang_x && text_time, UNIMP_E_NULL, "text_time: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0 && T_out > 0 && (use_cached || T_out == T), UNIMP_E_SHAPE,
                  "text_time: bad shape B=%d T=%d T_out=%d", B, T, T_out);
  const int wpb = 4;
  text_time_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
      lang_x, media_token_id, B, T, use_cached, T_out, text_time);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_mask_labels(const int64_t* input_ids, int64_t answer_id, int64_t endofchunk_id,
                                 int64_t media_id, int64_t pad_id, int64_t* labels, int B, int T,
                                 void* stream) {
  UNIMP_CHECK_ARG(input_ids && labels, UNIMP_E_NULL, "mask_labels: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0, UNIMP_E_SHAPE, "mask_labels: bad shape");
  const int wpb = 4;
  mask_labels_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
      input_ids, answer_id, endofchunk_id, media_id, pad_id, labels, B, T);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_sumsq(const void* grad, int64_t n, float* acc, int dtype, void* stream) {
  UNIMP_CHECK_ARG(grad && acc, UNIMP_E_NULL, "sumsq: NULL pointer");
  UNIMP_CHECK_ARG(aligned16(grad), UNIMP_E_ALIGN, "sumsq: grad must be 16-byte aligned");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "sumsq: dtype");
  if (n <= 0) return 0;
  const int npv = dtype == UNIMP_BF16 ? 8 : 4;
  int64_t blocks = (n / npv + 255) / 256;
  if (blocks > 8 * UNIMP_NUM_SMS) blocks = 8 * UNIMP_NUM_SMS;
  if (blocks < 1) blocks = 1;
  if (dtype == UNIMP_BF16)
    sumsq_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)grad, n, acc);
  else
    sumsq_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)grad, n,
                                                                             acc);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_adamw_step(float* master, void* param, const void* grad, float* exp_avg,
                                float* exp_avg_sq, int64_t n, const float* hyper, float beta1,
                                float beta2, float eps, float weight_decay, const float* gnorm_sq,
                                float max_norm, float grad_scale, int background, int dtype,
                                void* stream) {
  UNIMP_CHECK_ARG(master && param && grad && exp_avg && exp_avg_sq && hyper, UNIMP_E_NULL,
                  "adamw_step: NULL pointer");
  UNIMP_CHECK_ARG(aligned16(master) && aligned16(exp_avg) && aligned16(exp_avg_sq), UNIMP_E_ALIGN,
                  "adamw_step: fp32 state must be 16-byte aligned");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "adamw_step: dtype");
  if (n <= 0) return 0;
  int64_t blocks = (n / 4 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  if (background) {
    // Meant to run UNDER other kernels (beside the frozen ViT forward): a grid the block scheduler
    // dispatches at once (2 CTAs of 256 threads per SM, ~25 % of an SM's thread slots and registers),
    // so that the CTAs of the concurrently running compute-bound kernels fit next to it; 4 vectors
    // per thread in flight keep the HBM pipe full from that small footprint.
    if (blocks > 2 * UNIMP_NUM_SMS) blocks = 2 * UNIMP_NUM_SMS;
    // the kernels it runs beside (cuBLAS GEMMs, the attention kernels) want the largest shared-memory
    // carve-out; an SM cannot host CTAs with different carve-outs at once, so ask for the same one
    static bool carve = false;
    if (!carve) {
      cudaFuncSetAttribute(adamw_kernel<__nv_bfloat16, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(adamw_kernel<float, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      carve = true;
    }
    if (dtype == UNIMP_BF16)
      adamw_kernel<__nv_bfloat16, 4><<<(unsigned)blocks, 256, 0, st>>>(
          master, (__nv_bfloat16*)param, (const __nv_bfloat16*)grad, exp_avg, exp_avg_sq, n, hyper, beta1,
          beta2, eps, weight_decay, gnorm_sq, max_norm, grad_scale);
    else
      adamw_kernel<float, 4><<<(unsigned)blocks, 256, 0, st>>>(
          master, (float*)param, (const float*)grad, exp_avg, exp_avg_sq, n, hyper, beta1, beta2, eps,
          weight_decay, gnorm_sq, max_norm, grad_scale);
    UNIMP_CHECK_LAUNCH();
    return 0;
  }
  if (blocks > 16 * UNIMP_NUM_SMS) blocks = 16 * UNIMP_NUM_SMS;
  if (blocks < 1) blocks = 1;
  if (dtype == UNIMP_BF16)
    adamw_kernel<__nv_bfloat16, 1><<<(unsigned)blocks, 256, 0, st>>>(
        master, (__nv_bfloat16*)param, (const __nv_bfloat16*)grad, exp_avg, exp_avg_sq, n, hyper,
        beta1, beta2, eps, weight_decay, gnorm_sq, max_norm, grad_scale);
  else
    adamw_kernel<float, 1><<<(unsigned)blocks, 256, 0, st>>>(
        master, (float*)param, (const float*)grad, exp_avg, exp_avg_sq, n, hyper, beta1, beta2, eps,
        weight_decay, gnorm_sq, max_norm, grad_scale);
  UNIMP_CHECK_LAUNCH();
  return 0;
}
