"""torch.autograd wrappers over the C ABI (include/unimp_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every op below launches
hand-written sm_100a kernels through ctypes.  No CPU path, no fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import BF16, F32, View, check

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _dt(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise _lib.UnimpError("unimp_b200 ops need CUDA tensors (there is no CPU fallback)")
    try:
        return _DT[t.dtype]
    except KeyError:
        raise _lib.UnimpError(f"unsupported dtype {t.dtype} (float32 / bfloat16 only)") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _view3(t: torch.Tensor) -> View:
    """(B, L, H*dh) tensor with unit inner stride -> unimp_view_t (strides in elements)."""
    assert t.dim() == 3 and t.stride(2) == 1, (t.shape, t.stride())
    return View(t.data_ptr(), t.stride(0), t.stride(1))


def _inner_contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.stride(-1) == 1 else t.contiguous()


# --------------------------------------------------------------------------- text_time

def text_time(lang_x: torch.Tensor, media_token_id: int, *, use_cached: bool = False,
              T_out: int | None = None) -> torch.Tensor:
    """int32 (B, T_out): cumsum of `lang_x == media_token_id` (or its total, broadcast, when
    `use_cached`) — MaskedCrossAttention's text_time."""
    assert lang_x.dtype == torch.int64 and lang_x.dim() == 2 and lang_x.is_cuda
    lang_x = lang_x.contiguous()
    B, T = lang_x.shape
    T_out = T if T_out is None else T_out
    out = torch.empty((B, T_out), dtype=torch.int32, device=lang_x.device)
    check(_lib.load().unimp_text_time(lang_x.data_ptr(), int(media_token_id), B, T,
                                      int(use_cached), T_out, out.data_ptr(), _stream()),
          "unimp_text_time")
    return out


def mask_labels(input_ids, *, answer_token_id, endofchunk_token_id, media_token_id, pad_token_id):
    """GPU form of reference UniMP/mmrec.py:143-168."""
    assert input_ids.dtype == torch.int64 and input_ids.dim() == 2 and input_ids.is_cuda
    input_ids = input_ids.contiguous()
    B, T = input_ids.shape
    out = torch.empty_like(input_ids)
    check(_lib.load().unimp_mask_labels(input_ids.data_ptr(), int(answer_token_id),
                                        int(endofchunk_token_id), int(media_token_id),
                                        int(pad_token_id), out.data_ptr(), B, T, _stream()),
          "unimp_mask_labels")
    return out


# --------------------------------------------------------------------------- attention

class _Attention(torch.autograd.Function):
    """q (B,Lq,H*64); kv (B,Lk,2*H*64) packed [k | v]; text_time int32 (B,Lq) or None."""

    @staticmethod
    def forward(ctx, q, kv, tt, heads, n, Ti, scale, force_simt):
        dt = _dt(q)
        assert kv.dtype == q.dtype
        q = _inner_contig(q)
        kv = _inner_contig(kv)
        B, Lq, inner = q.shape
        Lk = kv.shape[1]
        dh = inner // heads
        assert kv.shape[2] == 2 * inner and kv.shape[0] == B
        k, v = kv[..., :inner], kv[..., inner:]
        o = torch.empty((B, Lq, inner), dtype=q.dtype, device=q.device)
        lse = torch.empty((B, heads, Lq), dtype=torch.float32, device=q.device)
        lib = _lib.load()
        if tt is not None:
            assert tt.dtype == torch.int32 and tt.shape == (B, Lq) and tt.is_contiguous()
        if force_simt:
            rc = lib.unimp__attn_fwd_simt(_view3(q), _view3(k), _view3(v), _ptr(tt), _view3(o),
                                          lse.data_ptr(), B, Lq, Lk, heads, n, Ti, dh, scale, dt,
                                          _stream())
        elif tt is not None:
            assert Lk == Ti * n
            rc = lib.unimp_xattn_fwd(_view3(q), _view3(k), _view3(v), tt.data_ptr(), _view3(o),
                                     lse.data_ptr(), B, Lq, Ti, n, heads, dh, scale, dt, _stream())
        else:
            rc = lib.unimp_attn_fwd(_view3(q), _view3(k), _view3(v), _view3(o), lse.data_ptr(), B,
                                    Lq, Lk, heads, dh, scale, dt, _stream())
        check(rc, "attention forward")
        ctx.save_for_backward(q, kv, o, lse, tt)
        ctx.cfg = (heads, n, Ti, scale, force_simt, dt)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, kv, o, lse, tt = ctx.saved_tensors
        heads, n, Ti, scale, force_simt, dt = ctx.cfg
        B, Lq, inner = q.shape
        Lk = kv.shape[1]
        dh = inner // heads
        d_o = _inner_contig(d_o)
        k, v = kv[..., :inner], kv[..., inner:]
        dq = torch.empty_like(q, memory_format=torch.contiguous_format)
        dkv = torch.empty((B, Lk, 2 * inner), dtype=kv.dtype, device=kv.device)
        dk, dv = dkv[..., :inner], dkv[..., inner:]
        lib = _lib.load()
        ws = torch.empty(lib.unimp_attn_bwd_workspace(B, Lq, Lk, heads, dh), dtype=torch.uint8,
                         device=q.device)
        if force_simt:
            rc = lib.unimp__attn_bwd_simt(_view3(q), _view3(k), _view3(v), _ptr(tt), _view3(o),
                                          _view3(d_o), lse.data_ptr(), ws.data_ptr(), _view3(dq),
                                          _view3(dk), _view3(dv), B, Lq, Lk, heads, n, Ti, dh,
                                          scale, dt, _stream())
        elif tt is not None:
            rc = lib.unimp_xattn_bwd(_view3(q), _view3(k), _view3(v), tt.data_ptr(), _view3(o),
                                     _view3(d_o), lse.data_ptr(), ws.data_ptr(), _view3(dq),
                                     _view3(dk), _view3(dv), B, Lq, Ti, n, heads, dh, scale, dt,
                                     _stream())
        else:
            rc = lib.unimp_attn_bwd(_view3(q), _view3(k), _view3(v), _view3(o), _view3(d_o),
                                    lse.data_ptr(), ws.data_ptr(), _view3(dq), _view3(dk),
                                    _view3(dv), B, Lq, Lk, heads, dh, scale, dt, _stream())
        check(rc, "attention backward")
        return dq, dkv, None, None, None, None, None, None


def masked_cross_attention(q, kv, text_time_i32, *, heads: int, n_latents: int, scale: float,
                           force_simt: bool = False):
    """K1: each query row attends the `n_latents` keys of image text_time-1 (0 -> zero row)."""
    Lk = kv.shape[1]
    assert Lk % n_latents == 0
    return _Attention.apply(q, kv, text_time_i32, heads, n_latents, Lk // n_latents, float(scale),
                            force_simt)


def attention(q, kv, *, heads: int, scale: float, force_simt: bool = False):
    """K2/K3: unmasked softmax(scale * q k^T) v, packed kv."""
    return _Attention.apply(q, kv, None, heads, kv.shape[1], 1, float(scale), force_simt)


def _weight_grad(w, g2, x2):
    """dW = g2^T x2 for a bias-free Linear weight: straight into the optimizer's flat gradient
    buffer when the parameter is registered for direct accumulation (returns None), else returned
    to autograd."""
    if getattr(w, "_unimp_direct", False) and w.grad is not None:
        if w._unimp_fresh:
            torch.mm(g2.t(), x2, out=w.grad)
            w._unimp_fresh = False
        else:
            w.grad.addmm_(g2.t(), x2)
        _grad_ready(w)
        return None
    return g2.t() @ x2


class _XattnBlock(torch.autograd.Function):
    """y = to_out( masked_attention( to_q(x_ln), to_kv(media) ) ) in ONE kernel
    (`unimp_xattn_block_fwd`); backward = the unfused pieces (dense projections on cuBLAS,
    `unimp_xattn_bwd` for the core) on the q / o / lse the forward saved."""

    @staticmethod
    def forward(ctx, x_ln, w_q, kv, tt, w_out, heads, n, scale):
        dt = _dt(x_ln)
        x_ln = x_ln.contiguous()
        kv = _inner_contig(kv)
        B, T, D = x_ln.shape
        inner = w_q.shape[0]
        dh = inner // heads
        Lk = kv.shape[1]
        Ti = Lk // n
        assert w_q.shape == (inner, D) and w_out.shape == (D, inner) and kv.shape == (B, Lk, 2 * inner)
        assert w_q.is_contiguous() and w_out.is_contiguous() and w_q.dtype == x_ln.dtype == kv.dtype
        assert tt.dtype == torch.int32 and tt.shape == (B, T) and tt.is_contiguous()
        k, v = kv[..., :inner], kv[..., inner:]
        q = torch.empty((B, T, inner), dtype=x_ln.dtype, device=x_ln.device)
        o = torch.empty_like(q)
        lse = torch.empty((B, heads, T), dtype=torch.float32, device=x_ln.device)
        y = torch.empty((B, T, D), dtype=x_ln.dtype, device=x_ln.device)
        check(_lib.load().unimp_xattn_block_fwd(x_ln.data_ptr(), w_q.data_ptr(), _view3(k), _view3(v),
                                                tt.data_ptr(), w_out.data_ptr(), q.data_ptr(),
                                                o.data_ptr(), lse.data_ptr(), y.data_ptr(), B, T, Ti, n,
                                                heads, dh, D, float(scale), dt, _stream()),
              "unimp_xattn_block_fwd")
        ctx.save_for_backward(x_ln, w_q, kv, tt, w_out, q, o, lse)
        ctx.cfg = (heads, n, Ti, float(scale), dt)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_ln, w_q, kv, tt, w_out, q, o, lse = ctx.saved_tensors
        heads, n, Ti, scale, dt = ctx.cfg
        B, T, D = x_ln.shape
        inner = q.shape[2]
        dh = inner // heads
        Lk = kv.shape[1]
        dy = dy.contiguous()
        dy2 = dy.reshape(B * T, D)
        d_o = (dy2 @ w_out).view(B, T, inner)                      # y = o @ w_out^T
        d_wout = _weight_grad(w_out, dy2, o.reshape(B * T, inner)) if ctx.needs_input_grad[4] else None
        k, v = kv[..., :inner], kv[..., inner:]
        dq = torch.empty_like(q)
        dkv = torch.empty((B, Lk, 2 * inner), dtype=kv.dtype, device=kv.device)
        dk, dv = dkv[..., :inner], dkv[..., inner:]
        lib = _lib.load()
        ws = torch.empty(lib.unimp_attn_bwd_workspace(B, T, Lk, heads, dh), dtype=torch.uint8, device=q.device)
        check(lib.unimp_xattn_bwd(_view3(q), _view3(k), _view3(v), tt.data_ptr(), _view3(o), _view3(d_o),
                                  lse.data_ptr(), ws.data_ptr(), _view3(dq), _view3(dk), _view3(dv), B, T,
                                  Ti, n, heads, dh, scale, dt, _stream()), "unimp_xattn_bwd (fused block)")
        dq2 = dq.reshape(B * T, inner)
        d_x = (dq2 @ w_q).view(B, T, D) if ctx.needs_input_grad[0] else None     # q = x_ln @ w_q^T
        d_wq = _weight_grad(w_q, dq2, x_ln.reshape(B * T, D)) if ctx.needs_input_grad[1] else None
        return d_x, d_wq, (dkv if ctx.needs_input_grad[2] else None), None, d_wout, None, None, None


def xattn_block_supported(x_ln, kv, *, heads: int, n_latents: int) -> bool:
    if not x_ln.is_cuda or x_ln.dtype != torch.bfloat16:
        return False
    B, T, D = x_ln.shape
    inner = kv.shape[2] // 2
    return bool(_lib.load().unimp_xattn_block_supported(T, kv.shape[1] // n_latents, n_latents, heads,
                                                        inner // heads, D, BF16))


def xattn_block(x_ln, w_q, kv, text_time_i32, w_out, *, heads: int, n_latents: int, scale: float):
    """K1-fused: to_q -> masked media-located attention -> to_out as one cluster kernel.
    x_ln (B,T,D) = norm(x); w_q = to_q.weight (H*64, D); kv = to_kv(media) packed (B,Ti*n,2*H*64);
    w_out = to_out.weight (D, H*64).  Returns to_out(attention) (B,T,D)."""
    return _XattnBlock.apply(x_ln, w_q, kv, text_time_i32, w_out, heads, n_latents, float(scale))


def xattn_decode(q, kv, n_media_i32, *, heads: int, n_latents: int, scale: float):
    """a12: single-token decode against cached packed K/V. q (B,1,H*64)."""
    dt = _dt(q)
    q = _inner_contig(q)
    kv = _inner_contig(kv)
    B, one, inner = q.shape
    assert one == 1
    Lk = kv.shape[1]
    k, v = kv[..., :inner], kv[..., inner:]
    o = torch.empty_like(q, memory_format=torch.contiguous_format)
    check(_lib.load().unimp_xattn_decode(_view3(q), _view3(k), _view3(v), n_media_i32.data_ptr(),
                                         _view3(o), B, Lk // n_latents, n_latents, heads,
                                         inner // heads, float(scale), dt, _stream()),
          "unimp_xattn_decode")
    return o


def lm_decode_attention(qkv, cos, sin, k_cache, v_cache, indir, add_mask, cursor, *, heads: int,
                        head_dim: int, rotary_dim: int, scale: float):
    """a12 / f4: one GPT-NeoX self-attention decode step (`unimp_lm_decode_attn`): rotary on the new
    token's packed qkv (B,1,H*3*dh), K/V written in place at the device-side `cursor`, attention
    over cache positions 0..cursor through the beam indirection `indir` (B,Tmax) int32.
    `add_mask` (B,Tmax): 0 / -inf.  Returns (B,1,H*dh)."""
    dt = _dt(qkv)
    B = qkv.shape[0]
    Tmax = k_cache.shape[2]
    assert qkv.is_contiguous() and k_cache.is_contiguous() and v_cache.is_contiguous()
    assert k_cache.shape == (B, heads, Tmax, head_dim) and k_cache.dtype == qkv.dtype
    assert indir.dtype == torch.int32 and indir.shape == (B, Tmax) and indir.is_contiguous()
    assert add_mask.dtype == qkv.dtype and add_mask.numel() == B * Tmax and add_mask.is_contiguous()
    assert cursor.dtype == torch.int64 and cursor.is_cuda
    cos = cos.reshape(B, rotary_dim).contiguous()
    sin = sin.reshape(B, rotary_dim).contiguous()
    out = torch.empty((B, 1, heads * head_dim), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().unimp_lm_decode_attn(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), k_cache.data_ptr(),
                                           v_cache.data_ptr(), indir.data_ptr(), add_mask.data_ptr(),
                                           cursor.data_ptr(), out.data_ptr(), B, heads, Tmax, head_dim,
                                           rotary_dim, float(scale), dt, _stream()), "unimp_lm_decode_attn")
    return out


BEAM_TOPK = os.environ.get("UNIMP_BEAM_TOPK", "1") != "0"   # A/B switch: 0 = torch log_softmax + topk


def beam_topk(logits, running_scores, num_beams: int, k: int):
    """One beam-search step's candidate selection (`unimp_beam_topk`): logits (B*nb, V) fp32,
    running_scores (B, nb) fp32 -> (top_lp (B, k) sorted descending, top_idx (B, k) int64 over the
    flattened (nb * V) axis) — HF's log_softmax + running score + topk(2 * num_beams)."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.stride(1) == 1
    R, V = logits.shape
    B = R // num_beams
    running_scores = running_scores.reshape(-1).to(torch.float32).contiguous()
    lib = _lib.load()
    ws = torch.empty(lib.unimp_beam_topk_workspace(R, V), dtype=torch.uint8, device=logits.device)
    top_lp = torch.empty((B, k), dtype=torch.float32, device=logits.device)
    top_idx = torch.empty((B, k), dtype=torch.int64, device=logits.device)
    check(lib.unimp_beam_topk(logits.data_ptr(), logits.stride(0), running_scores.data_ptr(), B, num_beams, V, k,
                              ws.data_ptr(), top_lp.data_ptr(), top_idx.data_ptr(), _stream()), "unimp_beam_topk")
    return top_lp, top_idx


SMALL_M_LINEAR = os.environ.get("UNIMP_DECODE_GEMV", "1") != "0"   # A/B switch: 0 = cuBLAS for decode steps


def small_m_eligible(x, w) -> bool:
    """True if y = x w^T of this call runs on `unimp_linear_small_m`: inference, bf16, <= 8 rows."""
    K = x.shape[-1]
    return (SMALL_M_LINEAR and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.bfloat16
            and w.dtype == torch.bfloat16 and x.numel() // K <= 8 and K % 32 == 0 and w.is_contiguous()
            and w.data_ptr() % 16 == 0)


def linear_rows(x, w, b=None, act_gelu: bool = False):
    """Inference-time `F.linear(x, w, b)` (optionally followed by the exact GELU).  With <= 8 rows —
    the beams of one decode step — the weight is streamed once by `unimp_linear_small_m`
    (bias / GELU in its epilogue); otherwise cuBLAS + `gelu`."""
    if not small_m_eligible(x, w):
        y = torch.nn.functional.linear(x, w, b)
        return gelu(y) if act_gelu else y
    K = x.shape[-1]
    N = w.shape[0]
    x2 = x.reshape(-1, K)
    if not x2.is_contiguous() or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    y = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)
    check(_lib.load().unimp_linear_small_m(x2.data_ptr(), w.data_ptr(), _ptr(b), y.data_ptr(), x2.shape[0], N, K,
                                           1 if act_gelu else 0, _DT[x.dtype], _stream()), "unimp_linear_small_m")
    return y


# --------------------------------------------------------------------------- gate + residual + LN

class _GateResidualLN(torch.autograd.Function):
    """mode bits: 1 = has branch (gate+residual), 2 = has LayerNorm."""

    @staticmethod
    def forward(ctx, branch, x, gate, gamma, beta, eps):
        dt = _dt(x)
        beta_param = beta
        x = x.contiguous()
        D = x.shape[-1]
        rows = x.numel() // D
        has_b, has_ln = branch is not None, gamma is not None
        if has_b:
            branch = branch.contiguous()
            assert branch.shape == x.shape and branch.dtype == x.dtype
            assert gate is None or (gate.dtype == x.dtype and gate.numel() == 1)
        x_out = torch.empty_like(x) if has_b else None
        ln_out = torch.empty_like(x) if has_ln else None
        mean = rstd = None
        if has_ln:
            assert gamma.dtype == x.dtype and beta.dtype == x.dtype
            gamma, beta = gamma.contiguous(), beta.contiguous()
            mean = torch.empty(rows, dtype=torch.float32, device=x.device)
            rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        check(_lib.load().unimp_gate_residual_ln_fwd(
            _ptr(branch), x.data_ptr(), _ptr(gate), _ptr(gamma), _ptr(beta), _ptr(x_out),
            _ptr(ln_out), _ptr(mean), _ptr(rstd), rows, D, float(eps), dt, _stream()),
            "unimp_gate_residual_ln_fwd")
        ctx.save_for_backward(branch, x_out if has_b else x, gate, gamma, mean, rstd)
        ctx.cfg = (has_b, has_ln, rows, D, dt)
        ctx.beta_ref = beta_param   # the Parameter itself (its value is not needed: only its .grad)
        if has_b and has_ln:
            return x_out, ln_out
        return x_out if has_b else ln_out

    @staticmethod
    def backward(ctx, *grads):
        branch, xo, gate, gamma, mean, rstd = ctx.saved_tensors
        has_b, has_ln, rows, D, dt = ctx.cfg
        if has_b and has_ln:
            g_xout, g_ln = grads
        elif has_b:
            g_xout, g_ln = grads[0], None
        else:
            g_xout, g_ln = None, grads[0]
        if g_xout is not None:
            g_xout = g_xout.contiguous()
        if g_ln is not None:
            g_ln = g_ln.contiguous()
        lib = _lib.load()
        d_x = torch.empty_like(xo)
        # ungated residual (gate None): d_branch == d_x, the kernel neither reads `branch` nor
        # writes a second copy; the same tensor is handed to both inputs
        gated = has_b and gate is not None
        d_branch = torch.empty_like(xo) if gated else None
        need_gate = has_b and gate is not None and ctx.needs_input_grad[2]
        need_affine = has_ln and g_ln is not None and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        # Parameters registered for direct accumulation (FlatAdamW): the fold kernel writes (first
        # micro-batch of the step) or adds (later ones) straight into their flat-buffer gradients;
        # autograd gets None and the per-parameter `grad += d` launches disappear.
        beta = ctx.beta_ref
        accumulate = 0
        direct_gate = need_gate and _is_direct(gate)
        direct_affine = need_affine and _is_direct(gamma) and beta is not None and _is_direct(beta)
        if direct_gate:
            d_gate = gate.grad
            accumulate |= 0 if gate._unimp_fresh else 1
        else:
            d_gate = torch.empty_like(gate) if need_gate else None
        if direct_affine:
            assert gamma._unimp_fresh == beta._unimp_fresh
            d_gamma, d_beta = gamma.grad, beta.grad
            accumulate |= 0 if gamma._unimp_fresh else 2
        else:
            d_gamma = torch.empty_like(gamma) if need_affine else None   # frozen LN: no column sums
            d_beta = torch.empty_like(gamma) if need_affine else None
        ws = torch.empty(lib.unimp_gate_residual_ln_bwd_workspace(rows, D), dtype=torch.uint8,
                         device=xo.device)
        use_ln = has_ln and g_ln is not None
        check(lib.unimp_gate_residual_ln_bwd(
            _ptr(g_xout), _ptr(g_ln) if use_ln else None, _ptr(branch), xo.data_ptr(),
            _ptr(gate), _ptr(gamma) if use_ln else None, _ptr(mean), _ptr(rstd), d_x.data_ptr(),
            _ptr(d_branch), _ptr(d_gate), _ptr(d_gamma), _ptr(d_beta), ws.data_ptr(), rows, D,
            accumulate, dt, _stream()), "unimp_gate_residual_ln_bwd")
        if direct_gate:
            gate._unimp_fresh = False
            _grad_ready(gate)
            d_gate = None
        if direct_affine:
            gamma._unimp_fresh = beta._unimp_fresh = False
            _grad_ready(gamma)
            _grad_ready(beta)
            d_gamma = d_beta = None
        elif has_ln and d_gamma is None and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4]):
            d_gamma = torch.zeros_like(gamma)
            d_beta = torch.zeros_like(gamma)
        return (d_branch if gated else (d_x if has_b else None)), d_x, d_gate, d_gamma, d_beta, None


def gate_residual_ln(branch, x, gate, gamma, beta, eps: float = 1e-5):
    """K5: x_out = branch*tanh(gate) + x ; ln_out = LN(x_out)*gamma+beta -> (x_out, ln_out)."""
    return _GateResidualLN.apply(branch, x, gate, gamma, beta, eps)


def gate_residual(branch, x, gate):
    """K5 without the LayerNorm: branch*tanh(gate) + x (gate None: branch + x)."""
    return _GateResidualLN.apply(branch, x, gate, None, None, 0.0)


def layer_norm(x, gamma, beta, eps: float = 1e-5):
    """K5 without the gate: LN(x)*gamma+beta (MaskedCrossAttention.norm, perceiver norms)."""
    return _GateResidualLN.apply(None, x, None, gamma, beta, eps)


# --------------------------------------------------------------------------- focal CE

class _FocalCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weights, gamma, use_focal, group_size):
        dt = _dt(logits)
        assert logits.dim() == 3 and labels.shape == logits.shape[:2]
        B, T, V = logits.shape
        group_size = B if group_size is None else int(group_size)
        assert B % group_size == 0
        if logits.stride(2) != 1 or logits.stride(0) != T * logits.stride(1):
            logits = logits.contiguous()
        ld = logits.stride(1)
        labels = labels.to(device=logits.device, dtype=torch.int64).contiguous()
        weights = weights.to(device=logits.device, dtype=torch.float32).contiguous()
        dev = logits.device
        lib = _lib.load()
        row_lse = torch.empty(B * T, dtype=torch.float32, device=dev)
        row_pt = torch.empty(B * T, dtype=torch.float32, device=dev)
        acc = torch.empty(2 * (B // group_size), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.unimp_focal_ce_workspace(B, T, V, dt), dtype=torch.uint8, device=dev)
        check(lib.unimp_focal_ce_fwd(logits.data_ptr(), ld, labels.data_ptr(), weights.data_ptr(),
                                     float(gamma), int(use_focal), row_lse.data_ptr(),
                                     row_pt.data_ptr(), acc.data_ptr(), loss.data_ptr(),
                                     ws.data_ptr(), B, T, V, group_size, dt, _stream()),
              "unimp_focal_ce_fwd")
        ctx.save_for_backward(logits, labels, weights, row_lse, row_pt, acc)
        ctx.cfg = (float(gamma), int(use_focal), ld, dt, group_size)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        logits, labels, weights, row_lse, row_pt, acc = ctx.saved_tensors
        gamma, use_focal, ld, dt, group_size = ctx.cfg
        B, T, V = logits.shape
        g = g_loss.to(torch.float32).contiguous()
        buf = torch.empty((B, T, ld), dtype=logits.dtype, device=logits.device)
        check(_lib.load().unimp_focal_ce_bwd(logits.data_ptr(), ld, labels.data_ptr(),
                                             weights.data_ptr(), gamma, use_focal,
                                             row_lse.data_ptr(), row_pt.data_ptr(), acc.data_ptr(),
                                             g.data_ptr(), buf.data_ptr(), ld, B, T, V, group_size,
                                             dt, _stream()), "unimp_focal_ce_bwd")
        return (buf[..., :V] if ld != V else buf), None, None, None, None, None


def focal_ce(logits, labels, weights, *, gamma: float = 2.0, use_focal: bool = True,
             group_size: int | None = None):
    """K6: reference UniMP/mmrec.py:190-213 on (B,T,V) logits, (B,T) labels, (B,) weights.
    `group_size`: normalise per group of that many consecutive samples and average the groups
    (an accumulation window of micro-batches in one call); default: one group = the whole batch."""
    return _FocalCE.apply(logits, labels, weights, gamma, use_focal, group_size)


class _FocalCERows(torch.autograd.Function):
    """Same loss over pre-gathered logits rows (head + loss fusion)."""

    @staticmethod
    def forward(ctx, logits, targets, row_w, row_g, G, gamma, use_focal):
        dt = _dt(logits)
        assert logits.dim() == 2 and logits.stride(1) == 1
        R, V = logits.shape
        ld = logits.stride(0)
        dev = logits.device
        targets = targets.to(device=dev, dtype=torch.int64).contiguous()
        row_w = row_w.to(device=dev, dtype=torch.float32).contiguous()
        if row_g is not None:
            row_g = row_g.to(device=dev, dtype=torch.int32).contiguous()
        assert targets.shape == (R,) and row_w.shape == (R,)
        lib = _lib.load()
        row_lse = torch.empty(R, dtype=torch.float32, device=dev)
        row_pt = torch.empty(R, dtype=torch.float32, device=dev)
        acc = torch.empty(2 * G, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.unimp_focal_ce_workspace(R, 1, V, dt), dtype=torch.uint8, device=dev)
        check(lib.unimp_focal_ce_rows_fwd(logits.data_ptr(), ld, targets.data_ptr(), row_w.data_ptr(),
                                          _ptr(row_g), float(gamma), int(use_focal),
                                          row_lse.data_ptr(), row_pt.data_ptr(), acc.data_ptr(),
                                          loss.data_ptr(), ws.data_ptr(), R, int(G), V, dt, _stream()),
              "unimp_focal_ce_rows_fwd")
        ctx.save_for_backward(logits, targets, row_w, row_g, row_lse, row_pt, acc)
        ctx.cfg = (float(gamma), int(use_focal), ld, dt, int(G))
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        logits, targets, row_w, row_g, row_lse, row_pt, acc = ctx.saved_tensors
        gamma, use_focal, ld, dt, G = ctx.cfg
        R, V = logits.shape
        g = g_loss.to(torch.float32).contiguous()
        buf = torch.empty((R, ld), dtype=logits.dtype, device=logits.device)
        check(_lib.load().unimp_focal_ce_rows_bwd(logits.data_ptr(), ld, targets.data_ptr(),
                                                  row_w.data_ptr(), _ptr(row_g), gamma, use_focal,
                                                  row_lse.data_ptr(), row_pt.data_ptr(),
                                                  acc.data_ptr(), g.data_ptr(), buf.data_ptr(), ld,
                                                  R, G, V, dt, _stream()), "unimp_focal_ce_rows_bwd")
        return (buf[:, :V] if ld != V else buf), None, None, None, None, None, None


def focal_ce_rows(logits_rows, targets, row_weights, row_groups=None, *, n_groups: int = 1,
                  gamma: float = 2.0, use_focal: bool = True):
    """K6 on pre-gathered rows: logits (R,V), targets (R,) (-100 = unused slot), row_weights (R,),
    row_groups (R,) int32 in [0, n_groups).  loss = mean_g( sum_g(w*CE*focal) / n_valid_g )."""
    return _FocalCERows.apply(logits_rows, targets, row_weights, row_groups, n_groups, gamma, use_focal)


def gather_label_rows(labels, capacity=None):
    """Which (b,t) rows does the loss read?  Those whose SHIFTED label labels[b,t+1] != -100
    (reference UniMP/mmrec.py:194-195).  Returns (row_index (R,), targets (R,), overflow flag):
    `capacity=None`: exact R (one host sync); otherwise R = capacity, static shapes (CUDA-graph
    safe), unused slots have target -100 and point at row 0; `overflow` (0-dim bool tensor) is
    True if more rows were valid than `capacity` — the caller must make that loud."""
    B, T = labels.shape
    tgt = torch.full_like(labels, -100)
    tgt[:, :-1] = labels[:, 1:]
    tgt = tgt.reshape(-1)
    valid = tgt != -100
    if capacity is None:
        idx = valid.nonzero().squeeze(1)
        return idx, tgt[idx], None
    R = int(capacity)
    pos = valid.cumsum(0) - 1
    slot = torch.where(valid & (pos < R), pos, torch.full_like(pos, R))
    buf = torch.full((R + 1,), B * T, dtype=torch.int64, device=labels.device)
    buf.scatter_(0, slot, torch.arange(B * T, device=labels.device))
    idx = buf[:R]
    pad = idx >= B * T
    idx = idx.masked_fill(pad, 0)
    return idx, tgt[idx].masked_fill(pad, -100), valid.sum() > R


# --------------------------------------------------------------------------- LM / ViT fusions

class _RotaryQKV(torch.autograd.Function):
    """GPT-NeoX rotary on the packed qkv projection; returns q,k,v (B,H,T,dh) views of ONE packed
    output buffer (no chunk/cat temporaries)."""

    @staticmethod
    def forward(ctx, qkv, cos, sin, H, dh, rot):
        dt = _dt(qkv)
        B, T, _ = qkv.shape
        assert qkv.is_contiguous() and qkv.shape[2] == 3 * H * dh
        cos, sin = cos.contiguous(), sin.contiguous()
        assert cos.dtype == qkv.dtype and cos.shape[-1] == rot and cos.shape[-2] == T
        cb = cos.shape[0] if cos.dim() == 3 else 1
        cs_bs = T * rot if cb > 1 else 0
        out = torch.empty_like(qkv)
        check(_lib.load().unimp_rotary_qkv_fwd(qkv.data_ptr(), out.data_ptr(), cos.data_ptr(),
                                               sin.data_ptr(), B, T, H, dh, rot, cs_bs, dt,
                                               _stream()), "unimp_rotary_qkv_fwd")
        ctx.save_for_backward(cos, sin)
        ctx.cfg = (B, T, H, dh, rot, cs_bs, dt)
        v5 = out.view(B, T, H, 3, dh)
        q, k, v = (v5[:, :, :, i].transpose(1, 2) for i in range(3))
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        import ctypes
        cos, sin = ctx.saved_tensors
        B, T, H, dh, rot, cs_bs, dt = ctx.cfg
        gs = []
        for g in (dq, dk, dv):
            if g.stride(-1) != 1:
                g = g.contiguous()
            gs.append(g)
        strides = (ctypes.c_int64 * 9)(*[s for g in gs for s in (g.stride(0), g.stride(1), g.stride(2))])
        d_qkv = torch.empty((B, T, 3 * H * dh), dtype=gs[0].dtype, device=gs[0].device)
        check(_lib.load().unimp_rotary_qkv_bwd(gs[0].data_ptr(), gs[1].data_ptr(), gs[2].data_ptr(),
                                               strides, cos.data_ptr(), sin.data_ptr(),
                                               d_qkv.data_ptr(), B, T, H, dh, rot, cs_bs, dt,
                                               _stream()), "unimp_rotary_qkv_bwd")
        return d_qkv, None, None, None, None, None


def rotary_qkv(qkv, cos, sin, *, heads: int, head_dim: int, rotary_dim: int):
    """qkv (B,T,H*3*dh) projection output -> rotated q, k and v as (B,H,T,dh) views."""
    return _RotaryQKV.apply(qkv, cos, sin, heads, head_dim, rotary_dim)


class _LMAttention(torch.autograd.Function):
    """K4: causal (+ key padding) self-attention of a GPT-NeoX layer, head dim 80, bf16
    (`unimp_lm_attn_fwd/bwd`).  q,k,v: (B,H,T,dh) views of the rotated packed projection."""

    @staticmethod
    def forward(ctx, q, k, v, key_bits, scale):
        B, H, T, dh = q.shape
        assert q.stride() == k.stride() == v.stride() and q.stride(3) == 1
        o = torch.empty((B, T, H * dh), dtype=q.dtype, device=q.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=q.device)
        kb = key_bits.data_ptr() if key_bits is not None else None
        check(_lib.load().unimp_lm_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), q.stride(2),
                                            q.stride(1), kb, o.data_ptr(), lse.data_ptr(), B, T, H, dh,
                                            float(scale), _dt(q), _stream()), "unimp_lm_attn_fwd")
        ctx.save_for_backward(q, k, v, o, lse, key_bits)
        ctx.scale = float(scale)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse, key_bits = ctx.saved_tensors
        B, H, T, dh = q.shape
        d_o = d_o.contiguous()
        dq32 = torch.empty((B, H, (T + 127) // 128 * 128, 84), dtype=torch.float32, device=q.device)
        delta = torch.empty((B, H, T), dtype=torch.float32, device=q.device)
        dk = torch.empty((B, T, H, dh), dtype=q.dtype, device=q.device)
        dv = torch.empty_like(dk)
        kb = key_bits.data_ptr() if key_bits is not None else None
        check(_lib.load().unimp_lm_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), q.stride(2),
                                            q.stride(1), kb, o.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                            dq32.data_ptr(), delta.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                            B, T, H, dh,
                                            ctx.scale, _dt(q), _stream()), "unimp_lm_attn_bwd")
        return (dq32[:, :, :T, :dh].to(q.dtype), dk.transpose(1, 2), dv.transpose(1, 2), None, None)


class _RotaryLMAttention(torch.autograd.Function):
    """GPT-NeoX rotary + K4 attention as ONE autograd node: `unimp_rotary_qkv_fwd` -> `unimp_lm_attn_fwd`;
    backward `unimp_lm_attn_bwd` -> `unimp_rotary_qkv_bwd_f32q`, which reads the fp32 dq scratch
    directly (no separate fp32 -> bf16 pass, no per-tensor autograd hops)."""

    @staticmethod
    def forward(ctx, qkv, cos, sin, key_bits, H, dh, rot, scale):
        dt = _dt(qkv)
        B, T, _ = qkv.shape
        assert qkv.is_contiguous() and qkv.shape[2] == 3 * H * dh
        cos, sin = cos.contiguous(), sin.contiguous()
        assert cos.dtype == qkv.dtype and cos.shape[-1] == rot and cos.shape[-2] == T
        cs_bs = T * rot if (cos.dim() == 3 and cos.shape[0] > 1) else 0
        lib = _lib.load()
        rq = torch.empty_like(qkv)
        check(lib.unimp_rotary_qkv_fwd(qkv.data_ptr(), rq.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, T, H, dh,
                                       rot, cs_bs, dt, _stream()), "unimp_rotary_qkv_fwd")
        es = rq.element_size()
        o = torch.empty((B, T, H * dh), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        kb = key_bits.data_ptr() if key_bits is not None else None
        base = rq.data_ptr()
        check(lib.unimp_lm_attn_fwd(base, base + dh * es, base + 2 * dh * es, T * H * 3 * dh, H * 3 * dh, 3 * dh, kb,
                                    o.data_ptr(), lse.data_ptr(), B, T, H, dh, float(scale), dt, _stream()),
              "unimp_lm_attn_fwd")
        ctx.save_for_backward(cos, sin, rq, o, lse, key_bits)
        ctx.cfg = (B, T, H, dh, rot, cs_bs, dt, float(scale))
        return o

    @staticmethod
    def backward(ctx, d_o):
        import ctypes
        cos, sin, rq, o, lse, key_bits = ctx.saved_tensors
        B, T, H, dh, rot, cs_bs, dt, scale = ctx.cfg
        lib = _lib.load()
        d_o = d_o.contiguous()
        Tp = (T + 127) // 128 * 128
        dev = rq.device
        dq32 = torch.empty((B, H, Tp, 84), dtype=torch.float32, device=dev)
        delta = torch.empty((B, H, T), dtype=torch.float32, device=dev)
        dk = torch.empty((B, T, H, dh), dtype=rq.dtype, device=dev)
        dv = torch.empty_like(dk)
        kb = key_bits.data_ptr() if key_bits is not None else None
        es = rq.element_size()
        base = rq.data_ptr()
        check(lib.unimp_lm_attn_bwd(base, base + dh * es, base + 2 * dh * es, T * H * 3 * dh, H * 3 * dh, 3 * dh, kb,
                                    o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dq32.data_ptr(),
                                    delta.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, T, H, dh, scale, dt,
                                    _stream()), "unimp_lm_attn_bwd")
        # (b, h, t) strides in elements: dq in the fp32 scratch, dk / dv as (B,T,H,dh)
        strides = (ctypes.c_int64 * 9)(H * Tp * 84, Tp * 84, 84, T * H * dh, dh, H * dh, T * H * dh, dh, H * dh)
        d_qkv = torch.empty((B, T, 3 * H * dh), dtype=rq.dtype, device=dev)
        check(lib.unimp_rotary_qkv_bwd_f32q(dq32.data_ptr(), dk.data_ptr(), dv.data_ptr(), strides, cos.data_ptr(),
                                            sin.data_ptr(), d_qkv.data_ptr(), B, T, H, dh, rot, cs_bs, dt,
                                            _stream()), "unimp_rotary_qkv_bwd_f32q")
        return d_qkv, None, None, None, None, None, None, None


def rotary_lm_attention(qkv, cos, sin, key_bits=None, *, heads: int, head_dim: int, rotary_dim: int, scale: float):
    """qkv (B,T,H*3*dh) projection output -> rotary -> causal (+ key padding) attention -> (B,T,H*dh)."""
    return _RotaryLMAttention.apply(qkv, cos, sin, key_bits, heads, head_dim, rotary_dim, scale)


def lm_attention_supported_shape(x, T: int, H: int, dh: int) -> bool:
    return x.is_cuda and bool(_lib.load().unimp_lm_attn_supported(T, H, dh, _dt(x)))


def lm_attention_supported(q) -> bool:
    """bf16, head dim 80 (RedPajama-INCITE 3B: 32 heads x 80)."""
    B, H, T, dh = q.shape
    return q.is_cuda and bool(_lib.load().unimp_lm_attn_supported(T, H, dh, _dt(q)))


def key_bits(attention_mask):
    """(B,T) attention_mask (bool / uint8 / int64; nonzero = real token) -> (B, 2*ceil(T/64)) int32
    words of 32 keys for `lm_attention`."""
    m = attention_mask
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    elif m.dtype not in (torch.uint8, torch.int64):
        m = m.to(torch.int64)
    m = m.contiguous()
    B, T = m.shape
    bits = torch.empty((B, 2 * ((T + 63) // 64)), dtype=torch.int32, device=m.device)
    check(_lib.load().unimp_key_bits(m.data_ptr(), m.element_size(), bits.data_ptr(), B, T, _stream()),
          "unimp_key_bits")
    return bits


def lm_attention(q, k, v, key_bits=None, *, scale: float):
    """softmax(scale*q k^T + causal & key-padding mask) v; q,k,v (B,H,T,dh) strided views sharing one
    layout -> (B,T,H*dh) contiguous (what `GPTNeoXAttention.dense` consumes)."""
    return _LMAttention.apply(q, k, v, key_bits, scale)


def quick_gelu_(x):
    """In-place CLIP QuickGELU (no autograd: used inside the frozen, no_grad vision tower)."""
    assert x.is_contiguous()
    check(_lib.load().unimp_quick_gelu(x.data_ptr(), x.numel(), _dt(x), _stream()), "unimp_quick_gelu")
    return x


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(_lib.load().unimp_gelu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _dt(x), _stream()),
              "unimp_gelu_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        check(_lib.load().unimp_gelu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(),
                                         _dt(x), _stream()), "unimp_gelu_bwd")
        return dx


def gelu(x):
    """Exact (erf) GELU, `nn.GELU()` of upstream's FeedForward and GPT-NeoX `mlp.act`
    (`unimp_gelu_fwd/bwd`; numel must be a multiple of 8 (bf16) / 4 (fp32))."""
    return _Gelu.apply(x)


# --------------------------------------------------------------------------- direct grad accumulation

def _is_direct(p) -> bool:
    return getattr(p, "_unimp_direct", False) and p.grad is not None


def _grad_ready(w):
    h = getattr(w, "_unimp_grad_ready", None)
    if h is not None:
        h(w)


class _LinearAcc(torch.autograd.Function):
    """Bias-free linear whose weight gradient is written by the dW GEMM itself into the
    optimizer's flat gradient buffer (`w.grad`, a view): beta = 0 on the first micro-batch of a
    step, beta = 1 afterwards.  Removes autograd's separate `grad += dW` pass over 1.15 B
    parameters per micro-batch (and the memset of their share of the buffer)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.nn.functional.linear(x, w)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dx = g @ w if ctx.needs_input_grad[0] else None
        if not ctx.needs_input_grad[1]:
            return dx, None
        g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])
        if getattr(w, "_unimp_direct", False) and w.grad is not None:
            if w._unimp_fresh:
                torch.mm(g2.t(), x2, out=w.grad)
                w._unimp_fresh = False
            else:
                w.grad.addmm_(g2.t(), x2)
            _grad_ready(w)
            return dx, None
        return dx, g2.t() @ x2


def linear_acc(x, w):
    """F.linear(x, w) (no bias) with direct gradient accumulation when the optimizer enabled it."""
    return _LinearAcc.apply(x, w)


class _EmbeddingAcc(torch.autograd.Function):
    """Embedding lookup whose backward scatter-adds the B*T gradient rows straight into the flat
    gradient buffer instead of materialising a dense (V, D) gradient and adding it."""

    @staticmethod
    def forward(ctx, ids, w):
        ctx.save_for_backward(ids, w)
        return torch.nn.functional.embedding(ids, w)

    @staticmethod
    def backward(ctx, g):
        ids, w = ctx.saved_tensors
        if getattr(w, "_unimp_direct", False) and w.grad is not None:
            if w._unimp_fresh:
                w.grad.zero_()
                w._unimp_fresh = False
            w.grad.index_add_(0, ids.reshape(-1), g.reshape(-1, g.shape[-1]))
            _grad_ready(w)
            return None, None
        dw = torch.zeros_like(w)
        dw.index_add_(0, ids.reshape(-1), g.reshape(-1, g.shape[-1]))
        return None, dw


def embedding_acc(ids, w):
    return _EmbeddingAcc.apply(ids, w)


# --------------------------------------------------------------------------- optimizer pieces

_WEIGHTS_EPOCH = [0]


def weights_epoch() -> int:
    """Counts optimizer steps taken through raw-pointer kernels (they do not bump a tensor's
    `_version`); caches derived from trainable weights key on it."""
    return _WEIGHTS_EPOCH[0]


def bump_weights_epoch():
    _WEIGHTS_EPOCH[0] += 1


def sumsq_(grad: torch.Tensor, acc: torch.Tensor):
    """acc[0] += sum(grad^2)."""
    check(_lib.load().unimp_sumsq(grad.data_ptr(), grad.numel(), acc.data_ptr(), _dt(grad),
                                  _stream()), "unimp_sumsq")


def adamw_step_(master, param, grad, exp_avg, exp_avg_sq, *, hyper, beta1, beta2, eps, weight_decay,
                gnorm_sq=None, max_norm=0.0, grad_scale=1.0, background=False):
    """hyper: device float32[3] = (lr, 1-beta1^t, sqrt(1-beta2^t)); see adamw_hyper().
    `background`: launch geometry for running under other kernels on a low-priority stream."""
    check(_lib.load().unimp_adamw_step(master.data_ptr(), param.data_ptr(), grad.data_ptr(),
                                       exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
                                       hyper.data_ptr(), float(beta1), float(beta2), float(eps),
                                       float(weight_decay), _ptr(gnorm_sq), float(max_norm),
                                       float(grad_scale), int(background), _dt(param), _stream()),
          "unimp_adamw_step")


def adamw_hyper(lr: float, beta1: float, beta2: float, step: int):
    """Host-side values of the `hyper` vector for optimizer step `step` (1-based)."""
    return [float(lr), 1.0 - beta1 ** step, (1.0 - beta2 ** step) ** 0.5]
