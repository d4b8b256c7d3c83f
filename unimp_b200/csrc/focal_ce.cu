// K6 — task-weighted focal cross-entropy head (reference UniMP/mmrec.py:190-213).
//
// HBM-bound.  Forward reads each VALID logits row exactly once (rows whose shifted label
// is -100 are never touched), split over V in chunks so that a handful of valid rows still
// spreads over many SMs; a single-CTA finish pass combines the per-chunk (max, sumexp)
// partials in a fixed order, so the loss is bit-reproducible run to run.
// Algorithmic bytes: fwd N_valid*V*sizeof(T); bwd N_valid*V*sizeof(T) read + B*T*V*sizeof(T)
// written (autograd needs the dense d_logits).
//
// Rows are NOT assumed 16-byte aligned (V = 74 053 is odd): each row is cut into 16-byte
// "slots" relative to its first aligned address; the ragged head/tail slots go scalar.
#include "common.cuh"

namespace unimp {

constexpr int CE_THREADS = 256;
constexpr int CE_SLOTS_PER_THREAD = 4;  // 4 x 16 B in flight per thread

template <typename T>
struct RowSlots {
  // slot 0 = [0, head) (may be empty), slot j>=1 = [head+(j-1)*N, head+j*N) clipped to V
  static constexpr int N = Vec16<T>::N;
  int head;
  int nslots;
  __device__ RowSlots(const T* row, int V) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(row);
    int h = (int)(((16 - (a & 15)) & 15) / sizeof(T));
    head = h < V ? h : V;
    nslots = 1 + (V - head + N - 1) / N;
  }
  __device__ __forceinline__ int begin(int j) const { return j == 0 ? 0 : head + (j - 1) * N; }
  __device__ __forceinline__ int end(int j, int V) const {
    int e = j == 0 ? head : head + j * N;
    return e < V ? e : V;
  }
};

template <typename T>
__device__ __forceinline__ void load_slot(const T* row, int V, const RowSlots<T>& rs, int j,
                                          float* f, float fill) {
  constexpr int N = Vec16<T>::N;
  const int b = rs.begin(j), e = rs.end(j, V);
  if (j > 0 && e - b == N) {
    Vec16<T> v;
    v.load_stream(row + b);
    v.unpack(f);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = (b + i < e) ? Elem<T>::to_f(row[b + i]) : fill;
  }
}

// Two row addressings share every kernel below:
//   dense  (T_ > 0): logits (B,T,V); row (b,t) is scored against labels[b,t+1]; weight weights[b];
//                    normalisation group b / group_size.
//   rows   (T_ == 0): logits (R,V) are pre-gathered rows (head + loss fusion: only rows with a
//                    valid label ever reach the head GEMM); labels[r] is the row's own target
//                    (-100 = padding slot), weights[r] its sample's weight, groups[r] its group.
struct RowMeta {
  int64_t y;
  float w;
  int g;
};
__device__ __forceinline__ int64_t row_target(const int64_t* __restrict__ labels, int row, int T_) {
  if (T_ > 0) return (row % T_ == T_ - 1) ? -100 : labels[row + 1];
  return labels[row];
}
__device__ __forceinline__ RowMeta row_meta(const int64_t* __restrict__ labels,
                                            const float* __restrict__ weights,
                                            const int32_t* __restrict__ groups, int row, int T_,
                                            int group_size) {
  RowMeta m;
  m.y = row_target(labels, row, T_);
  if (T_ > 0) {
    m.w = weights[row / T_];
    m.g = (row / T_) / group_size;
  } else {
    m.w = weights[row];
    m.g = groups ? groups[row] : 0;
  }
  return m;
}

// grid (rows, nchunks). partial[(row*nchunks + c)*2 + {0,1}] = (max, sum exp(x-max)).
template <typename T>
__global__ void __launch_bounds__(CE_THREADS)
focal_ce_partial_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                        float* __restrict__ partial, int T_, int V) {
  constexpr int N = Vec16<T>::N;
  const int row = blockIdx.x;  // b*T + t (dense) or gathered row index
  if (row_target(labels, row, T_) == -100) return;
  const T* rp = logits + (int64_t)row * ld;
  RowSlots<T> rs(rp, V);
  const int slots_per_cta = CE_THREADS * CE_SLOTS_PER_THREAD;
  const int s0 = blockIdx.y * slots_per_cta;
  float f[CE_SLOTS_PER_THREAD][N];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < CE_SLOTS_PER_THREAD; ++k) {
    const int j = s0 + k * CE_THREADS + threadIdx.x;
    if (j < rs.nslots) {
      load_slot(rp, V, rs, j, f[k], -INFINITY);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) f[k][i] = -INFINITY;
    }
  }
#pragma unroll
  for (int k = 0; k < CE_SLOTS_PER_THREAD; ++k)
#pragma unroll
    for (int i = 0; i < N; ++i) m = fmaxf(m, f[k][i]);
  __shared__ float sh[32];
  m = block_max(m, sh);
  float s = 0.f;
  if (m > -INFINITY) {
    const float ml2 = m * 1.4426950408889634f;
#pragma unroll
    for (int k = 0; k < CE_SLOTS_PER_THREAD; ++k)
#pragma unroll
      for (int i = 0; i < N; ++i) s += exp2f(f[k][i] * 1.4426950408889634f - ml2);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    float* p = partial + ((int64_t)row * gridDim.y + blockIdx.y) * 2;
    p[0] = m;
    p[1] = s;
  }
}

// Single CTA, fixed summation order => deterministic loss.
template <typename T>
__global__ void __launch_bounds__(1024)
focal_ce_finish_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                       const float* __restrict__ weights, const int32_t* __restrict__ groups,
                       const float* __restrict__ partial,
                       int nchunks, float gamma, int use_focal, float* __restrict__ row_lse,
                       float* __restrict__ row_pt, float* __restrict__ acc, float* __restrict__ loss,
                       int n_rows, int G, int T_, int group_size) {
  // Samples are normalised in groups of `group_size` (one group = one micro-batch of the
  // reference's accumulation window): loss = mean_g( sum_g(w*CE*focal) / n_valid_g ).
  float total = 0.f;
  __shared__ float sh[32];
  for (int g = 0; g < G; ++g) {
  float lsum = 0.f, nval = 0.f;
  // dense: group g owns a contiguous row range; rows mode: filter by the row's group id
  const int row_lo = T_ > 0 ? g * group_size * T_ : 0;
  const int rows = T_ > 0 ? (g + 1) * group_size * T_ : n_rows;
  for (int row = row_lo + threadIdx.x; row < rows; row += blockDim.x) {
    const RowMeta rm = row_meta(labels, weights, groups, row, T_, group_size);
    const int64_t y = rm.y;
    if (y == -100 || rm.g != g) continue;
    const float* p = partial + (int64_t)row * nchunks * 2;
    float m = -INFINITY;
    for (int c = 0; c < nchunks; ++c) m = fmaxf(m, p[2 * c]);
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += p[2 * c + 1] * __expf(p[2 * c] - m);
    const float lse = m + logf(s);
    const float xy = Elem<T>::to_f(logits[(int64_t)row * ld + y]);
    const float ce = lse - xy;
    const float pt = expf(xy - lse);
    row_lse[row] = lse;
    row_pt[row] = pt;
    float l = rm.w * ce;
    if (use_focal) l *= powf(fmaxf(1.f - pt, 0.f), gamma);
    lsum += l;
    nval += 1.f;
  }
  lsum = block_sum(lsum, sh);
  nval = block_sum(nval, sh);
  if (threadIdx.x == 0) {
    acc[2 * g] = lsum;
    acc[2 * g + 1] = nval;
  }
  total += lsum / nval;  // NaN when nothing is valid, as the reference (mmrec.py:213)
  }
  if (threadIdx.x == 0) *loss = total / G;
}

// grid (B*T, nchunks): writes every element of d_logits.
template <typename T>
__global__ void __launch_bounds__(CE_THREADS)
focal_ce_bwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                    const float* __restrict__ weights, const int32_t* __restrict__ groups,
                    float gamma, int use_focal,
                    const float* __restrict__ row_lse, const float* __restrict__ row_pt,
                    const float* __restrict__ acc, const float* __restrict__ g_loss,
                    T* __restrict__ d_logits, int64_t ld_out, int T_, int V, int group_size, int G) {
  constexpr int N = Vec16<T>::N;
  const int row = blockIdx.x;
  T* op = d_logits + (int64_t)row * ld_out;
  RowSlots<T> os(op, V);
  const int slots_per_cta = CE_THREADS * CE_SLOTS_PER_THREAD;
  const int s0 = blockIdx.y * slots_per_cta;
  const RowMeta rm = row_meta(labels, weights, groups, row, T_, group_size);
  const int64_t y = rm.y;
  float coef = 0.f, lse = 0.f;
  const T* rp = logits + (int64_t)row * ld;
  bool vec_in = false;
  if (y != -100) {
    lse = row_lse[row];
    const float pt = row_pt[row];
    float c = 1.f;
    if (use_focal) {
      const float omp = fmaxf(1.f - pt, 0.f);
      const float ce = -logf(fmaxf(pt, 1e-38f));
      c = powf(omp, gamma) + gamma * powf(omp, gamma - 1.f) * pt * ce;
    }
    coef = c * rm.w * (*g_loss) / (acc[2 * rm.g + 1] * G);
    vec_in = ((reinterpret_cast<uintptr_t>(rp) ^ reinterpret_cast<uintptr_t>(op)) & 15) == 0;
  }
#pragma unroll
  for (int k = 0; k < CE_SLOTS_PER_THREAD; ++k) {
    const int j = s0 + k * CE_THREADS + threadIdx.x;
    if (j >= os.nslots) continue;
    const int b = os.begin(j), e = os.end(j, V);
    float f[N];
    if (y == -100) {
#pragma unroll
      for (int i = 0; i < N; ++i) f[i] = 0.f;
    } else {
      if (vec_in && j > 0 && e - b == N) {
        Vec16<T> v;
        v.load_stream(rp + b);
        v.unpack(f);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = (b + i < e) ? Elem<T>::to_f(rp[b + i]) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float p = __expf(f[i] - lse);
        if ((int64_t)(b + i) == y) p -= 1.f;
        f[i] = coef * p;
      }
    }
    if (j > 0 && e - b == N) {
      Vec16<T> v;
      v.pack(f);
      v.store(op + b);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (b + i < e) op[b + i] = Elem<T>::from_f(f[i]);
    }
  }
}

static inline int ce_nchunks(int V, int dtype) {
  const int n = dtype == UNIMP_BF16 ? 8 : 4;
  const int nslots = 2 + (V + n - 1) / n;  // upper bound incl. ragged head
  const int per = CE_THREADS * CE_SLOTS_PER_THREAD;
  return (nslots + per - 1) / per;
}

}  // namespace unimp

using namespace unimp;

extern "C" int64_t unimp_focal_ce_workspace(int B, int T, int V, int dtype) {
  return (int64_t)B * T * ce_nchunks(V, dtype) * 2 * sizeof(float);
}


static int focal_fwd_launch(const void* logits, int64_t ld, const int64_t* labels,
                            const float* weights, const int32_t* groups, float gamma, int use_focal,
                            float* row_lse, float* row_pt, float* acc, float* loss, void* workspace,
                            int n_rows, int G, int T, int V, int group_size, int dtype,
                            cudaStream_t st) {
  const int nch = ce_nchunks(V, dtype);
  dim3 grid(n_rows, nch);
  float* partial = (float*)workspace;
  if (dtype == UNIMP_BF16) {
    focal_ce_partial_kernel<__nv_bfloat16><<<grid, CE_THREADS, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, labels, partial, T, V);
    UNIMP_CHECK_LAUNCH();
    focal_ce_finish_kernel<__nv_bfloat16><<<1, 1024, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, labels, weights, groups, partial, nch, gamma, use_focal,
        row_lse, row_pt, acc, loss, n_rows, G, T, group_size);
  } else {
    focal_ce_partial_kernel<float><<<grid, CE_THREADS, 0, st>>>((const float*)logits, ld, labels,
                                                                  partial, T, V);
    UNIMP_CHECK_LAUNCH();
    focal_ce_finish_kernel<float><<<1, 1024, 0, st>>>((const float*)logits, ld, labels, weights,
                                                       groups, partial, nch, gamma, use_focal,
                                                       row_lse, row_pt, acc, loss, n_rows, G, T,
                                                       group_size);
  }
  UNIMP_CHECK_LAUNCH();
  return 0;
}

static int focal_bwd_launch(const void* logits, int64_t ld, const int64_t* labels,
                            const float* weights, const int32_t* groups, float gamma, int use_focal,
                            const float* row_lse, const float* row_pt, const float* acc,
                            const float* g_loss, void* d_logits, int64_t ld_out, int n_rows, int G,
                            int T, int V, int group_size, int dtype, cudaStream_t st) {
  dim3 grid(n_rows, ce_nchunks(V, dtype));
  if (dtype == UNIMP_BF16)
    focal_ce_bwd_kernel<__nv_bfloat16><<<grid, CE_THREADS, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, labels, weights, groups, gamma, use_focal, row_lse, row_pt,
        acc, g_loss, (__nv_bfloat16*)d_logits, ld_out, T, V, group_size, G);
  else
    focal_ce_bwd_kernel<float><<<grid, CE_THREADS, 0, st>>>(
        (const float*)logits, ld, labels, weights, groups, gamma, use_focal, row_lse, row_pt, acc,
        g_loss, (float*)d_logits, ld_out, T, V, group_size, G);
  UNIMP_CHECK_LAUNCH();
  return 0;
}

extern "C" int unimp_focal_ce_fwd(const void* logits, int64_t ld, const int64_t* labels,
                                  const float* weights, float gamma, int use_focal,
                                  float* row_lse, float* row_pt, float* acc, float* loss,
                                  void* workspace, int B, int T, int V, int group_size, int dtype,
                                  void* stream) {
  UNIMP_CHECK_ARG(logits && labels && weights && row_lse && row_pt && acc && loss && workspace,
                  UNIMP_E_NULL, "focal_ce_fwd: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 1 && V > 0 && ld >= V, UNIMP_E_SHAPE,
                  "focal_ce_fwd: bad shape B=%d T=%d V=%d ld=%lld", B, T, V, (long long)ld);
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "focal_ce_fwd: dtype");
  UNIMP_CHECK_ARG(group_size > 0 && B % group_size == 0, UNIMP_E_SHAPE,
                  "focal_ce_fwd: group_size=%d must divide B=%d", group_size, B);
  return focal_fwd_launch(logits, ld, labels, weights, nullptr, gamma, use_focal, row_lse, row_pt, acc,
                          loss, workspace, B * T, B / group_size, T, V, group_size, dtype,
                          (cudaStream_t)stream);
}

extern "C" int unimp_focal_ce_bwd(const void* logits, int64_t ld, const int64_t* labels,
                                  const float* weights, float gamma, int use_focal,
                                  const float* row_lse, const float* row_pt, const float* acc,
                                  const float* g_loss, void* d_logits, int64_t ld_out, int B, int T,
                                  int V, int group_size, int dtype, void* stream) {
  UNIMP_CHECK_ARG(logits && labels && weights && row_lse && row_pt && acc && g_loss && d_logits,
                  UNIMP_E_NULL, "focal_ce_bwd: NULL pointer");
  UNIMP_CHECK_ARG(B > 0 && T > 1 && V > 0 && ld >= V && ld_out >= V, UNIMP_E_SHAPE,
                  "focal_ce_bwd: bad shape");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "focal_ce_bwd: dtype");
  UNIMP_CHECK_ARG(group_size > 0 && B % group_size == 0, UNIMP_E_SHAPE,
                  "focal_ce_bwd: group_size=%d must divide B=%d", group_size, B);
  return focal_bwd_launch(logits, ld, labels, weights, nullptr, gamma, use_focal, row_lse, row_pt, acc,
                          g_loss, d_logits, ld_out, B * T, B / group_size, T, V, group_size, dtype,
                          (cudaStream_t)stream);
}

extern "C" int unimp_focal_ce_rows_fwd(const void* logits, int64_t ld, const int64_t* targets,
                                       const float* row_weights, const int32_t* row_groups,
                                       float gamma, int use_focal, float* row_lse, float* row_pt,
                                       float* acc, float* loss, void* workspace, int R, int G, int V,
                                       int dtype, void* stream) {
  UNIMP_CHECK_ARG(logits && targets && row_weights && row_lse && row_pt && acc && loss && workspace,
                  UNIMP_E_NULL, "focal_ce_rows_fwd: NULL pointer");
  UNIMP_CHECK_ARG(R > 0 && G > 0 && V > 0 && ld >= V && (G == 1 || row_groups), UNIMP_E_SHAPE,
                  "focal_ce_rows_fwd: bad shape R=%d G=%d V=%d ld=%lld", R, G, V, (long long)ld);
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "focal_ce_rows_fwd: dtype");
  return focal_fwd_launch(logits, ld, targets, row_weights, row_groups, gamma, use_focal, row_lse,
                          row_pt, acc, loss, workspace, R, G, 0, V, 1, dtype, (cudaStream_t)stream);
}

extern "C" int unimp_focal_ce_rows_bwd(const void* logits, int64_t ld, const int64_t* targets,
                                       const float* row_weights, const int32_t* row_groups,
                                       float gamma, int use_focal, const float* row_lse,
                                       const float* row_pt, const float* acc, const float* g_loss,
                                       void* d_logits, int64_t ld_out, int R, int G, int V, int dtype,
                                       void* stream) {
  UNIMP_CHECK_ARG(logits && targets && row_weights && row_lse && row_pt && acc && g_loss && d_logits,
                  UNIMP_E_NULL, "focal_ce_rows_bwd: NULL pointer");
  UNIMP_CHECK_ARG(R > 0 && G > 0 && V > 0 && ld >= V && ld_out >= V && (G == 1 || row_groups),
                  UNIMP_E_SHAPE, "focal_ce_rows_bwd: bad shape");
  UNIMP_CHECK_ARG(dtype == UNIMP_F32 || dtype == UNIMP_BF16, UNIMP_E_DTYPE, "focal_ce_rows_bwd: dtype");
  return focal_bwd_launch(logits, ld, targets, row_weights, row_groups, gamma, use_focal, row_lse,
                          row_pt, acc, g_loss, d_logits, ld_out, R, G, 0, V, 1, dtype,
                          (cudaStream_t)stream);
}
