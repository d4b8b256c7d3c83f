/*
 * unimp_b200.h — C ABI of the B200-native UniMP / OpenFlamingo hot path.
 *
 * The reference (weitianxin/UniMP) has NO native code and NO FFI: every GPU kernel it
 * runs is reached through PyTorch modules of the un-vendored `open_flamingo==2.0.1`
 * package (reference requirements.txt:35, imported at UniMP/mmrec.py:20-22) and through
 * the in-tree loss arithmetic (UniMP/mmrec.py:190-213).  Each entry point below names
 * the Python interface it replaces.  SURVEY.md §8(b) is the contract this header meets.
 *
 * Conventions (all functions):
 *   - every pointer is a DEVICE pointer unless stated; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises or allocates;
 *   - returns 0 on success, <0 for an invalid argument (UNIMP_E_*), >0 = cudaError_t;
 *     unimp_last_error_string() describes the last failure on the calling thread;
 *   - `dtype` is the storage type of activation tensors: accumulation, softmax and
 *     LayerNorm statistics are always fp32;
 *   - re-entrant and thread-safe; safe to capture in a CUDA graph.
 *   - compiled for sm_100a only.  There is no CPU path.
 */
#ifndef UNIMP_B200_H_
#define UNIMP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIMP_ABI_VERSION 1

typedef enum { UNIMP_F32 = 0, UNIMP_BF16 = 1 } unimp_dtype_t;

enum {
  UNIMP_E_NULL = -1,      /* required pointer is NULL */
  UNIMP_E_SHAPE = -2,     /* unsupported / inconsistent shape */
  UNIMP_E_DTYPE = -3,     /* unsupported dtype */
  UNIMP_E_ALIGN = -4,     /* pointer or stride not 16-byte aligned */
  UNIMP_E_DEVICE = -5     /* not running on an sm_100 device */
};

/* Strided (batch, row, head, d) view; d is contiguous, head stride = dh elements.
 * Element (b, r, h, d) lives at base + b*batch_stride + r*row_stride + h*dh + d
 * (strides in ELEMENTS).  This is what `to_q(x)` / `to_kv(media).chunk(2)` /
 * `in_proj(x).chunk(3)` produce without any copy. */
typedef struct {
  const void* ptr;
  int64_t batch_stride;
  int64_t row_stride;
} unimp_view_t;

typedef struct {
  void* ptr;
  int64_t batch_stride;
  int64_t row_stride;
} unimp_mview_t;

int unimp_version(void);
const char* unimp_last_error_string(void);
/* 1 if the current device is sm_100 (B200), else 0. Host-side query, no stream. */
int unimp_device_ok(void);

/* ---- a4: media_locations -> text_time ------------------------------------------------
 * Replaces `FlamingoLMMixin.forward`'s `media_locations = input_ids == media_token_id`
 * plus MaskedCrossAttention's `text_time = media_locations.cumsum(-1)` (or, with
 * use_cached != 0, `count_nonzero(media_locations)` of the CACHED prompt broadcast over
 * the T_out new tokens) — SURVEY §9; call site UniMP/mmrec.py:177-181.
 * lang_x (B,T) int64 -> text_time (B,T_out) int32.  T_out = T unless use_cached. */
int unimp_text_time(const int64_t* lang_x, int64_t media_token_id, int B, int T,
                    int use_cached, int T_out, int32_t* text_time, void* stream);

/* ---- a7 / K1: masked, media-located cross-attention core ------------------------------
 * Replaces the einsum/masked_fill/amax/softmax/masked_fill/einsum core of
 * `MaskedCrossAttention.forward`, only_attend_immediate_media=True.
 * q (B,T,H,dh) pre-projection-scaled NOT required: `scale` is applied in-kernel.
 * k,v (B,Ti*n,H,dh).  Row i of sample b attends the n keys of image text_time[b,i]-1;
 * text_time==0 -> output row 0, lse = -inf;  text_time > Ti -> uniform attention over
 * all Ti*n keys (upstream's all-masked softmax; kept for parity).
 * o (B,T,H,dh) ; lse (B,H,T) fp32 = log-sum-exp of scale*q.k over the attended keys.
 * dh must be 64; n (latents per image) must be 64 for the tensor-core path. */
int unimp_xattn_fwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* text_time,
                    unimp_mview_t o, float* lse, int B, int T, int Ti, int n, int H, int dh,
                    float scale, int dtype, void* stream);
/* Backward.  workspace: caller scratch of unimp_attn_bwd_workspace(B,T,Ti*n,H,dh) bytes,
 * 16-byte aligned.  dq/dk/dv are fully written (zero where nothing attends). */
int64_t unimp_attn_bwd_workspace(int Bt, int Lq, int Lk, int H, int dh);
int unimp_xattn_bwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* text_time,
                    unimp_view_t o, unimp_view_t d_o, const float* lse, void* workspace,
                    unimp_mview_t dq, unimp_mview_t dk, unimp_mview_t dv, int B, int T, int Ti,
                    int n, int H, int dh, float scale, int dtype, void* stream);

/* Test hooks (CUDA-core implementation forced; tt may be NULL = unmasked). */
int unimp__attn_fwd_simt(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                         unimp_mview_t o, float* lse, int B, int Lq, int Lk, int H, int n, int Ti,
                         int dh, float scale, int dtype, void* stream);
int unimp__attn_bwd_simt(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* tt,
                         unimp_view_t o, unimp_view_t d_o, const float* lse, void* workspace,
                         unimp_mview_t dq, unimp_mview_t dk, unimp_mview_t dv, int B, int Lq,
                         int Lk, int H, int n, int Ti, int dh, float scale, int dtype,
                         void* stream);

/* ---- K2 / K3: unmasked attention core (Perceiver 64x320, ViT-L/14 257x257) ------------
 * Replaces `PerceiverAttention.forward`'s einsum/softmax/einsum and the
 * ViT `nn.MultiheadAttention` core (the dead UniMP/xformers_model/clip.py:130-136 is
 * the same shape).  q (Bt,Lq,H,dh), k/v (Bt,Lk,H,dh) -> o (Bt,Lq,H,dh), lse (Bt,H,Lq). */
int unimp_attn_fwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_mview_t o, float* lse,
                   int Bt, int Lq, int Lk, int H, int dh, float scale, int dtype, void* stream);
int unimp_attn_bwd(unimp_view_t q, unimp_view_t k, unimp_view_t v, unimp_view_t o,
                   unimp_view_t d_o, const float* lse, void* workspace, unimp_mview_t dq,
                   unimp_mview_t dk, unimp_mview_t dv, int Bt, int Lq, int Lk, int H, int dh,
                   float scale, int dtype, void* stream);

/* ---- a7 fused: MaskedCrossAttention (after its LayerNorm) as ONE kernel ------------------
 * Replaces, in upstream `MaskedCrossAttention.forward` (helpers.py; call site reference
 * UniMP/mmrec.py:177-181): q = to_q(x_ln); the masked attention core above; to_out(out).
 *   x_ln (B,T,D) contiguous = self.norm(x);  w_q (H*dh, D) = to_q.weight;  k,v (B,Ti*n,H,dh)
 *   views of to_kv(media);  w_out (D, H*dh) = to_out.weight;  text_time (B,T) int32.
 * Outputs: y (B,T,D) contiguous = to_out(attention);  q, o (B,T,H*dh) contiguous and lse (B,H,T):
 * the projected queries / attention output / log-sum-exp, saved for unimp_xattn_bwd.
 * One 8-CTA thread-block cluster per 128-row tile (one head per CTA, x_ln tile TMA-multicast);
 * needs H == 8, dh == 64, n == 64, D % 128 == 0, D <= 2560, bf16.
 * unimp_xattn_block_supported returns 1 if the shape is covered, else 0. */
int unimp_xattn_block_supported(int T, int Ti, int n, int H, int dh, int D, int dtype);
int unimp_xattn_block_fwd(const void* x_ln, const void* w_q, unimp_view_t k, unimp_view_t v,
                          const int32_t* text_time, const void* w_out, void* q, void* o, float* lse,
                          void* y, int B, int T, int Ti, int n, int H, int dh, int D, float scale,
                          int dtype, void* stream);

/* ---- a12: decode step against cached cross-attention K/V ------------------------------
 * One new token per sequence attends the LAST image's n cached keys.  Upstream
 * recomputes to_kv(media) every step; the cache is new here.
 * q (B,1,H,dh); k,v (B,Ti*n,H,dh) cached; n_media[b] = count of <image> in the prompt. */
int unimp_xattn_decode(unimp_view_t q, unimp_view_t k, unimp_view_t v, const int32_t* n_media,
                       unimp_mview_t o, int B, int Ti, int n, int H, int dh, float scale,
                       int dtype, void* stream);

/* ---- a12 / f4: the language model's side of a decode step (configs[3]) ------------------
 * Replaces, per GPT-NeoX layer and generated token, what HF runs inside `GenerationMixin.generate`
 * (reference call `UniMP/pipeline/eval/eval_exp.py:101-114`): apply_rotary_pos_emb on the new
 * token's q/k, `DynamicCache.update`, the attention over the cache, and — for beam search — the
 * `reorder_cache` copy of every layer's K/V.
 *   qkv (B, H, 3, dh): the packed query_key_value projection of the new token (HF layout);
 *   cos, sin (B, rot) of the token's position; k_cache, v_cache (B, H, Tmax, dh), written in place
 *   at slot *cursor (int64 on the device: capturable in a CUDA graph);
 *   indir (B, Tmax) int32: cache ROW that holds position t of beam b's history.  The kernel sets
 *   indir[b][*cursor] = b; a beam re-ordering step permutes the ROWS OF indir (B*Tmax ints), never
 *   the caches;  add_mask (B, Tmax) in `dtype`: 0 = visible, -inf = padding;
 *   out (B, H*dh) = softmax(q k^T * scale + add_mask) v over positions 0..*cursor.
 * dh in {32, 64, 80, 96, 128}; rot even, <= dh; Tmax <= 5120; fp32 or bf16. */
int unimp_lm_decode_attn(const void* qkv, const void* cos, const void* sin, void* k_cache,
                         void* v_cache, int32_t* indir, const void* add_mask, const int64_t* cursor,
                         void* out, int B, int H, int Tmax, int dh, int rot, float scale, int dtype,
                         void* stream);

/* Beam-search candidate selection of one decode step — what HF `GenerationMixin._beam_search` computes
 * with `log_softmax(logits)`, `+ running_beam_scores[:, :, None]` and `topk(2 * num_beams)` over the
 * flattened (beams * V) axis (reference call chain: `UniMP/pipeline/eval/eval_exp.py:101-114`).
 *   logits (B*nb, V) fp32, row stride ld; running (B*nb) fp32;
 *   top_lp (B, K) fp32 sorted descending, top_idx (B, K) int64 = beam_in_item * V + token.
 * K <= 16, K <= V, nb <= 64, nb * V < 2^31; ties break towards the smaller index.  workspace: device
 * buffer of unimp_beam_topk_workspace(B*nb, V) bytes. */
int64_t unimp_beam_topk_workspace(int rows, int V);
int unimp_beam_topk(const float* logits, int64_t ld, const float* running, int B, int nb, int V, int K,
                    void* workspace, float* top_lp, int64_t* top_idx, void* stream);

/* y (M, N) = act(x (M, K) . w (N, K)^T + bias (N) or NULL), M <= 8 rows (the beams of one decode
 * step): `nn.Linear` of GPT-NeoX / GatedCrossAttentionBlock at one token per sequence, as a
 * weight-streaming kernel (each weight is read once, 16-byte loads).  act: 0 none, 1 exact (erf)
 * GELU.  bf16 only; K % 32 == 0; x, w 16-byte aligned, rows contiguous. */
int unimp_linear_small_m(const void* x, const void* w, const void* bias, void* y, int M, int N, int K,
                         int act, int dtype, void* stream);

/* ---- a6 / K5: tanh-gate + residual + LayerNorm epilogue -------------------------------
 * Replaces `x = f(x) * tanh(gate) + x` of `GatedCrossAttentionBlock.forward` fused with
 * the LayerNorm that consumes it (the FF's LN, or MaskedCrossAttention.norm) — SURVEY §9.
 *   x_out = branch * tanh(*gate) + x        (branch == NULL: x_out = x, not written;
 *                                            gate is one element of `dtype`;
 *                                            gate == NULL: plain residual, factor 1)
 *   ln_out = LN(x_out) * gamma + beta       (gamma == NULL: LN skipped)
 * rows x D, D % 8 == 0, D <= 8192.  mean/rstd (rows) fp32 saved for backward. */
int unimp_gate_residual_ln_fwd(const void* branch, const void* x, const void* gate,
                               const void* gamma, const void* beta, void* x_out, void* ln_out,
                               float* mean, float* rstd, int64_t rows, int D, float eps,
                               int dtype, void* stream);
/* Backward.  g_xout / g_ln are the incoming grads of the two outputs (either may be NULL).
 * d_x = g_xout + LNbwd(g_ln);  d_branch = d_x * tanh(gate);
 * d_gate = sum(d_x * branch) * (1 - tanh^2)  ; d_gamma/d_beta = column sums (`dtype`; any of
 * the three may be NULL).  accumulate: bit 0 -> d_gate +=, bit 1 -> d_gamma / d_beta += instead of
 * = (the outputs are then the optimizer's gradient buffers and an earlier micro-batch of the
 * step has written them: no separate `grad += d` launches); 0 = write.
 * Ungated residual (branch != NULL, gate == NULL): d_branch == d_x, `branch` is not read and
 * d_branch may be NULL (the caller hands d_x to both inputs); if given it receives a copy.
 * partial is caller scratch of unimp_gate_residual_ln_bwd_workspace() bytes. */
int64_t unimp_gate_residual_ln_bwd_workspace(int64_t rows, int D);
int unimp_gate_residual_ln_bwd(const void* g_xout, const void* g_ln, const void* branch,
                               const void* x_out, const void* gate, const void* gamma,
                               const float* mean, const float* rstd, void* d_x, void* d_branch,
                               void* d_gate, void* d_gamma, void* d_beta, void* partial,
                               int64_t rows, int D, int accumulate, int dtype, void* stream);

/* ---- a10 / K6: task-weighted focal cross-entropy head ---------------------------------
 * Replaces reference UniMP/mmrec.py:190-213 (shift-by-one CE * weights * (1-pt)^gamma,
 * divided by the number of valid labels) in ONE read pass over the valid rows.
 * logits (B,T,V) with row stride `ld` elements (ld >= V); labels (B,T) int64 with -100;
 * row (b,t), t < T-1, is scored against labels[b,t+1].  weights (B) fp32.
 * use_focal == 0 reproduces plain weighted CE (reference --use_reweight off).
 * group_size: samples are normalised in groups of `group_size` consecutive samples (one group =
 * one micro-batch of the reference's gradient-accumulation window, UniMP/mmrec.py:175,213);
 * group_size == B is the reference's single-batch formula.  G = B / group_size.
 * Outputs: row_lse, row_pt (B*T) fp32 (undefined on ignored rows);
 *          acc[2g] = sum_{i in g} w*CE*(1-pt)^gamma, acc[2g+1] = n_valid_g (fp32, 2*G floats);
 *          *loss = mean_g(acc[2g]/acc[2g+1])  (NaN when a group has no valid label, as the
 *          reference does).
 * The sum runs in a fixed order: the loss is bit-reproducible run to run.
 * workspace: unimp_focal_ce_workspace(B,T,V,dtype) bytes of caller scratch. */
int64_t unimp_focal_ce_workspace(int B, int T, int V, int dtype);
int unimp_focal_ce_fwd(const void* logits, int64_t ld, const int64_t* labels,
                       const float* weights, float gamma, int use_focal, float* row_lse,
                       float* row_pt, float* acc, float* loss, void* workspace, int B, int T,
                       int V, int group_size, int dtype, void* stream);
/* d_logits (B,T,V) row stride ld_out, fully written: zero on ignored rows and at t = T-1.
 * g_loss: device scalar (fp32) upstream gradient of the loss. */
int unimp_focal_ce_bwd(const void* logits, int64_t ld, const int64_t* labels,
                       const float* weights, float gamma, int use_focal, const float* row_lse,
                       const float* row_pt, const float* acc, const float* g_loss,
                       void* d_logits, int64_t ld_out, int B, int T, int V, int group_size,
                       int dtype, void* stream);

/* Head + loss fusion: the same loss over PRE-GATHERED logits rows.  Only rows
 * (b,t) whose shifted label labels[b,t+1] != -100 matter to reference UniMP/mmrec.py:190-213
 * (24 of 1536 at configs[1]); the host gathers those hidden rows before the output head, so
 * logits is (R,V) with row stride ld.  targets[r] = the row's own label (-100 = unused padding
 * slot of a fixed-capacity gather), row_weights[r] = weights[b] of the row's sample,
 * row_groups[r] in [0,G) = its normalisation group (may be NULL when G == 1).
 * Same outputs as unimp_focal_ce_fwd; workspace: unimp_focal_ce_workspace(R, 1, V, dtype).
 * d_logits (R,V) row stride ld_out, fully written (zero rows for padding slots). */
int unimp_focal_ce_rows_fwd(const void* logits, int64_t ld, const int64_t* targets,
                            const float* row_weights, const int32_t* row_groups, float gamma,
                            int use_focal, float* row_lse, float* row_pt, float* acc, float* loss,
                            void* workspace, int R, int G, int V, int dtype, void* stream);
int unimp_focal_ce_rows_bwd(const void* logits, int64_t ld, const int64_t* targets,
                            const float* row_weights, const int32_t* row_groups, float gamma,
                            int use_focal, const float* row_lse, const float* row_pt,
                            const float* acc, const float* g_loss, void* d_logits, int64_t ld_out,
                            int R, int G, int V, int dtype, void* stream);

/* ---- a9 (f2): answer-span label masking on the GPU -----------------------------------
 * Replaces the Python double loop of reference UniMP/mmrec.py:143-168. */
int unimp_mask_labels(const int64_t* input_ids, int64_t answer_id, int64_t endofchunk_id,
                      int64_t media_id, int64_t pad_id, int64_t* labels, int B, int T,
                      void* stream);

/* ---- f3: elementwise fusions around the frozen towers -------------------------------
 * GPT-NeoX rotary embedding (HF `apply_rotary_pos_emb`, transformers gpt_neox) over the packed
 * (B,T,H,3*dh) output of `query_key_value` into `out` (same layout, out != qkv; v and the
 * non-rotary tail are copied); cos/sin are (1 or B, T, rot) in `dtype` (cs_batch_stride = 0 when
 * shared across the batch).  q/k/v then go to SDPA as strided views of `out`.  Backward un-rotates dq/dk and packs dq|dk|dv ((B,H,T,dh) views; strides9 is a HOST
 * array {q_sb,q_sh,q_st,k_sb,k_sh,k_st,v_sb,v_sh,v_st} in elements) into d_qkv (B,T,H,3*dh). */
int unimp_rotary_qkv_fwd(const void* qkv, void* out, const void* cos, const void* sin, int B, int T,
                         int H, int dh, int rot, int64_t cs_batch_stride, int dtype, void* stream);
int unimp_rotary_qkv_bwd(const void* dq, const void* dk, const void* dv, const int64_t* strides9,
                         const void* cos, const void* sin, void* d_qkv, int B, int T, int H, int dh,
                         int rot, int64_t cs_batch_stride, int dtype, void* stream);
/* ---- f3 / K4: causal self-attention of the GPT-NeoX decoder layers, head dim 80 -----------
 * Replaces the attention core of HF `GPTNeoXAttention.forward` (transformers gpt_neox, reached
 * through upstream `FlamingoLayer.forward` -> `self.decoder_layer(...)`; call site reference
 * UniMP/mmrec.py:177-181): softmax(scale * q k^T + mask) v per head, where key j is visible to
 * query i iff j <= i and bit j of key_bits[b] is set (what HF builds from a 2-D attention_mask;
 * key_bits == NULL: causal only).  q,k,v: (B,T,H,dh) views that share batch/row/head strides
 * (elements) — the rotated packed projection; o, d_o: (B,T,H*dh) contiguous; lse (B,H,T) fp32.
 * unimp_key_bits packs an attention_mask (B,T) of bool/uint8 (elem_size 1) or int64 (8),
 * nonzero = real token, into (B, 2*ceil(T/64)) words of 32 keys.
 * Backward: dq32 (B,H,Tp,84) fp32 scratch, Tp = T rounded up to 128 (zeroed and accumulated by
 * the call; dq = dq32[:, :, :T, :80]), delta (B,H,T) fp32 scratch (rowsum(dO o O), written by the
 * call), dk, dv (B,T,H,dh) in `dtype`, contiguous.  bf16, dh == 80 only (unimp_lm_attn_supported); rows that see no key
 * give o = 0. */
int unimp_lm_attn_supported(int T, int H, int dh, int dtype);
int unimp_key_bits(const void* mask, int elem_size, uint32_t* bits, int B, int T, void* stream);
int unimp_lm_attn_fwd(const void* q, const void* k, const void* v, int64_t batch_stride,
                      int64_t row_stride, int64_t head_stride, const uint32_t* key_bits, void* o,
                      float* lse, int B, int T, int H, int dh, float scale, int dtype, void* stream);
int unimp_lm_attn_bwd(const void* q, const void* k, const void* v, int64_t batch_stride,
                      int64_t row_stride, int64_t head_stride, const uint32_t* key_bits,
                      const void* o, const void* d_o, const float* lse, float* dq32, float* delta,
                      void* dk, void* dv, int B, int T, int H, int dh, float scale, int dtype,
                      void* stream);
/* Same with dq read as fp32 (the dq32 scratch of unimp_lm_attn_bwd, strides in fp32 elements): the
 * fp32 -> `dtype` conversion of dq happens in this pass instead of in one of its own. */
int unimp_rotary_qkv_bwd_f32q(const float* dq, const void* dk, const void* dv, const int64_t* strides9,
                              const void* cos, const void* sin, void* d_qkv, int B, int T, int H,
                              int dh, int rot, int64_t cs_batch_stride, int dtype, void* stream);
/* CLIP QuickGELU x*sigmoid(1.702x), in place (ViT MLP; forward only: the tower is frozen). */
int unimp_quick_gelu(void* x, int64_t n, int dtype, void* stream);
/* Exact (erf) GELU of the FeedForward blocks: open_flamingo helpers.FeedForward's nn.GELU()
 * (GatedCrossAttentionBlock.ff, PerceiverResampler ff) and GPT-NeoX mlp.act.
 * y = x * Phi(x);  dx = dy * (Phi(x) + x * phi(x)).  n elements, contiguous, n % (16/sizeof) == 0.
 * y may alias x; dx may alias dy.  bf16: Phi from a 1.5e-7-accurate rational erfc (well below bf16
 * resolution); fp32: erff. */
int unimp_gelu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream);
int unimp_gelu_bwd(const void* x, const void* dy, void* dx, int64_t n, int dtype, void* stream);

/* ---- f1: fused AdamW over a flat parameter group --------------------------------------
 * Replaces torch.optim.AdamW.step for one param group (reference UniMP/mmrec.py:671) with
 * grad-clip scaling folded in (clip coefficient = min(1, max_norm/(norm+1e-6)),
 * UniMP/mmrec.py:247-248).  master/m/v fp32, grad `dtype`; writes the `dtype` working
 * copy `param`.  gnorm_sq: device scalar, sum of squared grads (NULL = no clipping).
 * hyper: DEVICE float[3] = {lr, 1-beta1^t, sqrt(1-beta2^t)} — step-dependent scalars live in
 * device memory so that a captured CUDA graph of the step stays valid across steps.
 * background != 0: launch geometry for running UNDER other kernels (short-lived CTAs; the caller
 * puts it on a low-priority stream next to compute-bound work that does not read the parameters:
 * unimp_b200.train.GraphedTrainStep(defer_optimizer=True) overlaps it with the frozen ViT forward). */
int unimp_adamw_step(float* master, void* param, const void* grad, float* exp_avg,
                     float* exp_avg_sq, int64_t n, const float* hyper, float beta1, float beta2,
                     float eps, float weight_decay, const float* gnorm_sq, float max_norm,
                     float grad_scale, int background, int dtype,
                     void* stream);
/* acc[0] += sum(grad^2) (fp32). */
int unimp_sumsq(const void* grad, int64_t n, float* acc, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIMP_B200_H_ */
