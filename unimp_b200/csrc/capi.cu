// extern "C" surface of libunimp_b200.so — see include/unimp_b200.h for the contract.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace unimp {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("UNIMP_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

template <typename T>
int launch_attn_fwd_simt(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*, unimp_mview_t,
                         float*, int, int, int, int, int, int, float, cudaStream_t);
template <typename T>
int launch_attn_bwd_simt(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*, unimp_view_t,
                         unimp_view_t, const float*, void*, unimp_mview_t, unimp_mview_t,
                         unimp_mview_t, int, int, int, int, int, int, float, cudaStream_t);
template <typename T>
int launch_xattn_decode(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*, unimp_mview_t, int,
                        int, int, int, float, cudaStream_t);

// attn_tc.cu (tcgen05 + TMEM + TMA, bf16).  Returns UNIMP_E_SHAPE if the shape is not covered.
int launch_attn_fwd_tc(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                       unimp_mview_t o, float* lse, int B, int Lq, int Lk, int H, int n, int Ti,
                       float scale, cudaStream_t st);
const char* attn_fwd_tc_unsupported(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_mview_t o,
                                    const int32_t* tt, int Lq, int Lk, int n, int dh);
int launch_attn_bwd_tc(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                       unimp_view_t o, unimp_view_t d_o, const float* lse, void* workspace,
                       unimp_mview_t dq, unimp_mview_t dk, unimp_mview_t dv, int B, int Lq, int Lk,
                       int H, int n, int Ti, float scale, cudaStream_t st);
const char* attn_bwd_tc_unsupported(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_view_t d_o,
                                    unimp_mview_t dq, unimp_mview_t dk, unimp_mview_t dv,
                                    const int32_t* tt, int Lq, int Lk, int n, int dh);

// xattn_block.cu (one cluster kernel: to_q GEMM -> masked attention -> to_out GEMM)
const char* xattn_block_unsupported(int T, int Ti, int n, int H, int dh, int D, int dtype);
int launch_xattn_block_fwd(const void* x_ln, const void* w_q, unimp_view_t k, unimp_view_t v,
                           const int32_t* tt, const void* w_out, void* q, void* o, float* lse, void* y,
                           int B, int T, int Ti, int n, int D, float scale, cudaStream_t st);

static int check_attn_common(const char* who, unimp_view_t q, unimp_view_t k, unimp_view_t v,
                             const void* o, int B, int Lq, int Lk, int H, int dh, int dtype) {
  UNIMP_CHECK_ARG(q.ptr && k.ptr && v.ptr && o, UNIMP_E_NULL, "%s: NULL pointer", who);
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "%s: dtype %d", who,
                  dtype);
  UNIMP_CHECK_ARG(dh == 64, UNIMP_E_SHAPE, "%s: dim_head must be 64 (got %d)", who, dh);
  UNIMP_CHECK_ARG(B > 0 && Lq > 0 && Lk > 0 && H > 0, UNIMP_E_SHAPE,
                  "%s: bad shape B=%d Lq=%d Lk=%d H=%d", who, B, Lq, Lk, H);
  return 0;
}

}  // namespace unimp

using namespace unimp;

extern "C" int unimp_version(void) { return UNIMP_ABI_VERSION; }

extern "C" const char* unimp_last_error_string(void) { return g_err; }

extern "C" int unimp_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int64_t unimp_attn_bwd_workspace(int Bt, int Lq, int Lk, int H, int dh) {
  (void)Lk;
  const int64_t delta = (((int64_t)Bt * H * Lq + 3) / 4) * 4 * sizeof(float);
  const int64_t dq_acc = (int64_t)Bt * Lq * H * dh * sizeof(float);
  // tensor-core backward: fp32 dK/dV accumulators are not needed (one CTA owns a key block)
  return delta + dq_acc + 256;
}

static int attn_fwd_dispatch(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                             unimp_mview_t o, float* lse, int B, int Lq, int Lk, int H, int n, int Ti,
                             int dh, float scale, int dtype, int force_simt, cudaStream_t st) {
  // bf16 is the tensor-core path, full stop: a shape it does not cover is an ERROR, never a silent
  // detour through the CUDA cores.  The SIMT kernels serve fp32 (the 1e-4 parity mode) and the
  // explicit unimp__attn_*_simt test hooks only.
  if (dtype == UNIMP_BF16 && !force_simt) {
    const char* why = attn_fwd_tc_unsupported(q, k, v, o, tt, Lq, Lk, n, dh);
    UNIMP_CHECK_ARG(!why, UNIMP_E_SHAPE, "attention forward (bf16, tcgen05): %s [Lq=%d Lk=%d n=%d dh=%d]",
                    why, Lq, Lk, n, dh);
    return launch_attn_fwd_tc(q, k, v, tt, o, lse, B, Lq, Lk, H, n, Ti, scale, st);
  }
  if (dtype == UNIMP_BF16)
    return launch_attn_fwd_simt<__nv_bfloat16>(q, k, v, tt, o, lse, B, Lq, Lk, H, n, Ti, scale, st);
  return launch_attn_fwd_simt<float>(q, k, v, tt, o, lse, B, Lq, Lk, H, n, Ti, scale, st);
}

static int attn_bwd_dispatch(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                             unimp_view_t o, unimp_view_t d_o, const float* lse, void* ws,
                             unimp_mview_t dq, unimp_mview_t dk, unimp_mview_t dv, int B, int Lq,
                             int Lk, int H, int n, int Ti, int dh, float scale, int dtype,
                             int force_simt, cudaStream_t st) {
  if (dtype == UNIMP_BF16 && !force_simt) {
    const char* why = attn_bwd_tc_unsupported(q, k, v, d_o, dq, dk, dv, tt, Lq, Lk, n, dh);
    UNIMP_CHECK_ARG(!why, UNIMP_E_SHAPE, "attention backward (bf16, tcgen05): %s [Lq=%d Lk=%d n=%d dh=%d]",
                    why, Lq, Lk, n, dh);
    return launch_attn_bwd_tc(q, k, v, tt, o, d_o, lse, ws, dq, dk, dv, B, Lq, Lk, H, n, Ti, scale,
                              st);
  }
  if (dtype == UNIMP_BF16)
    return launch_attn_bwd_simt<__nv_bfloat16>(q, k, v, tt, o, d_o, lse, ws, dq, dk, dv, B, Lq, Lk, H,
                                               n, Ti, scale, st);
  return launch_attn_bwd_simt<float>(q, k, v, tt, o, d_o, lse, ws, dq, dk, dv, B, Lq, Lk, H, n, Ti,
                                     scale, st);
}

extern "C" int unimp_xattn_fwd(unimp_view_t q, unimp_view_t k, unimp_view_t v,
                               const int32_t* text_time, unimp_mview_t o, float* lse, int B, int T,
                               int Ti, int n, int H, int dh, float scale, int dtype, void* stream) {
  int rc = check_attn_common("xattn_fwd", q, k, v, o.ptr, B, T, Ti * n, H, dh, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(text_time && lse, UNIMP_E_NULL, "xattn_fwd: text_time/lse NULL");
  UNIMP_CHECK_ARG(Ti > 0 && n > 0, UNIMP_E_SHAPE, "xattn_fwd: Ti=%d n=%d", Ti, n);
  return attn_fwd_dispatch(q, k, v, text_time, o, lse, B, T, Ti * n, H, n, Ti, dh, scale, dtype, 0,
                           (cudaStream_t)stream);
}

extern "C" int unimp_xattn_bwd(unimp_view_t q, unimp_view_t k, unimp_view_t v,
                               const int32_t* text_time, unimp_view_t o, unimp_view_t d_o,
                               const float* lse, void* workspace, unimp_mview_t dq, unimp_mview_t dk,
                               unimp_mview_t dv, int B, int T, int Ti, int n, int H, int dh,
                               float scale, int dtype, void* stream) {
  int rc = check_attn_common("xattn_bwd", q, k, v, o.ptr, B, T, Ti * n, H, dh, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(text_time && lse && workspace && d_o.ptr && dq.ptr && dk.ptr && dv.ptr,
                  UNIMP_E_NULL, "xattn_bwd: NULL pointer");
  return attn_bwd_dispatch(q, k, v, text_time, o, d_o, lse, workspace, dq, dk, dv, B, T, Ti * n, H, n,
                           Ti, dh, scale, dtype, 0, (cudaStream_t)stream);
}

extern "C" int unimp_attn_fwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_mview_t o,
                              float* lse, int Bt, int Lq, int Lk, int H, int dh, float scale,
                              int dtype, void* stream) {
  int rc = check_attn_common("attn_fwd", q, k, v, o.ptr, Bt, Lq, Lk, H, dh, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(lse, UNIMP_E_NULL, "attn_fwd: lse NULL");
  return attn_fwd_dispatch(q, k, v, nullptr, o, lse, Bt, Lq, Lk, H, Lk, 1, dh, scale, dtype, 0,
                           (cudaStream_t)stream);
}

extern "C" int unimp_attn_bwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_view_t o,
                              unimp_view_t d_o, const float* lse, void* workspace, unimp_mview_t dq,
                              unimp_mview_t dk, unimp_mview_t dv, int Bt, int Lq, int Lk, int H,
                              int dh, float scale, int dtype, void* stream) {
  int rc = check_attn_common("attn_bwd", q, k, v, o.ptr, Bt, Lq, Lk, H, dh, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(lse && workspace && d_o.ptr && dq.ptr && dk.ptr && dv.ptr, UNIMP_E_NULL,
                  "attn_bwd: NULL pointer");
  return attn_bwd_dispatch(q, k, v, nullptr, o, d_o, lse, workspace, dq, dk, dv, Bt, Lq, Lk, H, Lk, 1,
                           dh, scale, dtype, 0, (cudaStream_t)stream);
}

extern "C" int unimp_xattn_block_supported(int T, int Ti, int n, int H, int dh, int D, int dtype) {
  return xattn_block_unsupported(T, Ti, n, H, dh, D, dtype) == nullptr ? 1 : 0;
}

extern "C" int unimp_xattn_block_fwd(const void* x_ln, const void* w_q, unimp_view_t k, unimp_view_t v,
                                     const int32_t* text_time, const void* w_out, void* q, void* o,
                                     float* lse, void* y, int B, int T, int Ti, int n, int H, int dh,
                                     int D, float scale, int dtype, void* stream) {
  UNIMP_CHECK_ARG(x_ln && w_q && k.ptr && v.ptr && text_time && w_out && q && o && lse && y, UNIMP_E_NULL,
                  "xattn_block_fwd: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 0 && Ti > 0, UNIMP_E_SHAPE, "xattn_block_fwd: bad shape B=%d T=%d Ti=%d", B, T,
                  Ti);
  const char* why = xattn_block_unsupported(T, Ti, n, H, dh, D, dtype);
  UNIMP_CHECK_ARG(!why, UNIMP_E_SHAPE, "xattn_block_fwd: %s [H=%d dh=%d n=%d D=%d]", why, H, dh, n, D);
  UNIMP_CHECK_ARG(aligned16(x_ln) && aligned16(w_q) && aligned16(w_out) && aligned16(q) && aligned16(o) &&
                      aligned16(y) && aligned16(k.ptr) && aligned16(v.ptr) && k.row_stride % 8 == 0 &&
                      v.row_stride % 8 == 0 && k.batch_stride % 8 == 0 && v.batch_stride % 8 == 0,
                  UNIMP_E_ALIGN, "xattn_block_fwd: pointers must be 16-byte aligned, strides multiples of 8");
  return launch_xattn_block_fwd(x_ln, w_q, k, v, text_time, w_out, q, o, lse, y, B, T, Ti, n, D, scale,
                                (cudaStream_t)stream);
}

// Test hooks: same contracts, CUDA-core implementation forced (tt may be NULL = unmasked).
extern "C" int unimp__attn_fwd_simt(unimp_view_t q, unimp_view_t k, unimp_view_t v,
                                    const int32_t* tt, unimp_mview_t o, float* lse, int B, int Lq,
                                    int Lk, int H, int n, int Ti, int dh, float scale, int dtype,
                                    void* stream) {
  int rc = check_attn_common("attn_fwd_simt", q, k, v, o.ptr, B, Lq, Lk, H, dh, dtype);
  if (rc) return rc;
  return attn_fwd_dispatch(q, k, v, tt, o, lse, B, Lq, Lk, H, n, Ti, dh, scale, dtype, 1,
                           (cudaStream_t)stream);
}
extern "C" int unimp__attn_bwd_simt(unimp_view_t q, unimp_view_t k, unimp_view_t v,
                                    const int32_t* tt, unimp_view_t o, unimp_view_t d_o,
                                    const float* lse, void* workspace, unimp_mview_t dq,
                                    unimp_mview_t dk, unimp_mview_t dv, int B, int Lq, int Lk, int H,
                                    int n, int Ti, int dh, float scale, int dtype, void* stream) {
  int rc = check_attn_common("attn_bwd_simt", q, k, v, o.ptr, B, Lq, Lk, H, dh, dtype);
  if (rc) return rc;
  return attn_bwd_dispatch(q, k, v, tt, o, d_o, lse, workspace, dq, dk, dv, B, Lq, Lk, H, n, Ti, dh,
                           scale, dtype, 1, (cudaStream_t)stream);
}

extern "C" int unimp_xattn_decode(unimp_view_t q, unimp_view_t k, unimp_view_t v,
                                  const int32_t* n_media, unimp_mview_t o, int B, int Ti, int n, int H,
                                  int dh, float scale, int dtype, void* stream) {
  int rc = check_attn_common("xattn_decode", q, k, v, o.ptr, B, 1, Ti * n, H, dh, dtype);
  if (rc) return rc;
  UNIMP_CHECK_ARG(n_media, UNIMP_E_NULL, "xattn_decode: n_media NULL");
  if (dtype == UNIMP_BF16)
    return launch_xattn_decode<__nv_bfloat16>(q, k, v, n_media, o, B, Ti, n, H, scale,
                                              (cudaStream_t)stream);
  return launch_xattn_decode<float>(q, k, v, n_media, o, B, Ti, n, H, scale, (cudaStream_t)stream);
}
