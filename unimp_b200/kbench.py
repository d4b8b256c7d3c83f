"""Isolated, cold-cache timings of the hand-written kernels at a workload's shapes.

Each kernel is launched over K independent input sets whose combined footprint exceeds the
126 MB L2 (so every launch reads cold data, no flush kernel in the timed region), the K launches
are captured in a CUDA graph (no Python/launch latency between them) and the graph replay is
timed with CUDA events on the launching stream: avg = elapsed / K.  Algorithmic bytes / FLOPs per
launch are the figures of DESIGN.md §4.
"""
from __future__ import annotations

import torch

from . import ops

L2_BYTES = 126 << 20


def _time_graph(fns, reps=3):
    """fns: list of zero-arg callables (one per input set). Returns average microseconds per call."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e3 / len(fns)
        best = t if best is None else min(best, t)
    del g
    return best


def _k(bytes_per_set, cap=48):
    return max(2, min(cap, (int(1.3 * L2_BYTES) + bytes_per_set - 1) // bytes_per_set))


def _eager_xattn(q, kv, tt, H, n, scale):
    """The reference's MaskedCrossAttention core as upstream executes it (dense, materialised
    (B,h,T,Ti*n) similarity + mask) — the eager-GPU column for K1."""
    B, T, inner = q.shape
    Lk = kv.shape[1]
    dh = inner // H
    k, v = kv.chunk(2, dim=-1)
    qh = q.view(B, T, H, dh).transpose(1, 2) * scale
    kh = k.reshape(B, Lk, H, dh).transpose(1, 2)
    vh = v.reshape(B, Lk, H, dh).transpose(1, 2)
    sim = torch.einsum("bhid,bhjd->bhij", qh, kh)
    media_time = (torch.arange(Lk // n, device=q.device) + 1).repeat_interleave(n)
    mask = tt[:, None, :, None] == media_time[None, None, None, :]
    sim = sim.masked_fill(~mask, -torch.finfo(sim.dtype).max)
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()
    attn = sim.softmax(dim=-1)
    attn = attn.masked_fill((tt == 0)[:, None, :, None], 0.0)
    out = torch.einsum("bhij,bhjd->bhid", attn, vh)
    return out.transpose(1, 2).reshape(B, T, inner)


def _eager_attn(q, kv, H, scale):
    """PerceiverAttention core as upstream executes it (einsum / amax / softmax / einsum)."""
    B, Lq, inner = q.shape
    Lk = kv.shape[1]
    dh = inner // H
    k, v = kv.chunk(2, dim=-1)
    qh = q.view(B, Lq, H, dh).transpose(1, 2) * scale
    kh = k.reshape(B, Lk, H, dh).transpose(1, 2)
    vh = v.reshape(B, Lk, H, dh).transpose(1, 2)
    sim = torch.einsum("bhid,bhjd->bhij", qh, kh)
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), vh)
    return out.transpose(1, 2).reshape(B, Lq, inner)


def _eager_focal(logits, labels, weights, gamma):
    """reference UniMP/mmrec.py:190-213, line for line, on the GPU."""
    n1, n2 = labels.shape[0], labels.shape[1] - 1
    shift_logits = logits[:, :-1, :].contiguous()
    lab = labels[:, 1:].contiguous().view(-1)
    shift_logits = shift_logits.view(-1, shift_logits.size(-1))
    lm_loss = torch.nn.functional.cross_entropy(shift_logits, lab, reduction="none").view(n1, n2)
    loss = (weights[:, None] * lm_loss).view(-1)
    p = torch.nn.functional.softmax(shift_logits, dim=-1)
    pt = p[torch.arange(len(shift_logits), device=p.device), lab]
    loss = loss * (1 - pt) ** gamma
    return loss.sum() / (lab != -100).sum()


def run(cfg, wl, peaks, *, dtype=torch.bfloat16, n_valid_rows=None, eager=True, only=None):
    """Returns {kernel: {avg_us, alg_bytes, flops, GB/s, TFLOP/s, frac_hbm, frac_tensor, sets}}.
    `eager`: also time what the reference would run on this GPU for the same op (plain PyTorch,
    same dtype, same shapes, same cold-cache method) as `eager_<kernel>` rows, and add
    `speedup_vs_eager` to our row."""
    out = {}

    def rec(name, us, byts, fl, sets):
        d = {"avg_us": us, "alg_bytes_per_launch": byts, "flops_per_launch": fl, "sets": sets,
             "GB/s": byts / us / 1e3, "frac_of_hbm_peak": byts / us / 1e3 / peaks["hbm_gbs"]}
        if fl:
            d["TFLOP/s"] = fl / us / 1e6
            d["frac_of_bf16_peak"] = d["TFLOP/s"] / peaks["bf16_tflops"]
        out[name] = d

    def want(section):
        return only is None or section in only

    torch.manual_seed(0)
    if want("xattn"):
        _run_xattn(cfg, wl, dtype, eager, rec)
    if want("vit") or want("perceiver"):
        _run_vit_perceiver(cfg, wl, dtype, eager, rec)
    if want("k5") or want("gelu"):
        _run_k5_gelu(cfg, wl, dtype, eager, rec)
    if want("lm"):
        _run_lm_attn(cfg, wl, dtype, eager, rec)
    if want("focal"):
        _run_focal(cfg, wl, dtype, eager, rec, out, n_valid_rows)
    if want("misc"):
        _run_misc(cfg, wl, dtype, rec)
    if only is not None and "decode" in only:
        _run_decode(cfg, dtype, eager, rec)
    if eager:
        for name in list(out):
            e = out.get("eager_" + name)
            if e is not None:
                out[name]["speedup_vs_eager"] = e["avg_us"] / out[name]["avg_us"]
        if "eager_vit_attn_fwd_sdpa" in out:
            out["vit_attn_fwd"]["speedup_vs_eager"] = out["eager_vit_attn_fwd_sdpa"]["avg_us"] / out["vit_attn_fwd"]["avg_us"]
    torch.cuda.empty_cache()
    return out


def _dims(cfg, wl, dtype):
    es = 2 if dtype == torch.bfloat16 else 4
    return ("cuda", es, wl.B, wl.T, wl.Ti, cfg.n_latents, cfg.xattn_heads, cfg.xattn_dim_head, cfg.lm_hidden,
            cfg.vocab)


def _run_xattn(cfg, wl, dtype, eager, rec):
    dev, es, B, T, Ti, n, H, dh, D, V = _dims(cfg, wl, dtype)
    inner = H * dh
    # ---- K1 masked x-attn ------------------------------------------------------------------
    xb = es * (2 * B * T * inner + 2 * B * Ti * n * inner) + 4 * B * T * H
    xf = 4.0 * H * dh * n * B * T
    K = _k(xb)
    loc = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        for j in range(Ti):
            loc[b, 1 + j * (T // Ti)] = True
    tt = loc.cumsum(-1).to(torch.int32).to(dev)
    qs = [torch.randn(B, T, inner, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    kvs = [torch.randn(B, Ti * n, 2 * inner, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    gos = [torch.randn(B, T, inner, device=dev, dtype=dtype) for _ in range(K)]
    fwd = [lambda q=q, kv=kv: ops.masked_cross_attention(q, kv, tt, heads=H, n_latents=n, scale=dh ** -0.5)
           for q, kv in zip(qs, kvs)]
    rec("xattn_fwd", _time_graph(fwd), xb, xf, K)
    os_ = [f() for f in fwd]
    bwd = [lambda o=o, q=q, kv=kv, g=g: torch.autograd.grad(o, (q, kv), g, retain_graph=True)
           for o, q, kv, g in zip(os_, qs, kvs, gos)]
    rec("xattn_bwd", _time_graph(bwd), 2.5 * xb, 2.5 * xf, K)
    if eager:
        efw = [lambda q=q, kv=kv: _eager_xattn(q, kv, tt, H, n, dh ** -0.5) for q, kv in zip(qs, kvs)]
        rec("eager_xattn_fwd", _time_graph(efw), xb, xf, K)
        eos = [f() for f in efw]
        rec("eager_xattn_bwd", _time_graph([lambda o=o, q=q, kv=kv, g=g: torch.autograd.grad(o, (q, kv), g, retain_graph=True)
                                            for o, q, kv, g in zip(eos, qs, kvs, gos)]), 2.5 * xb, 2.5 * xf, K)
        del efw, eos
    del qs, kvs, gos, os_, fwd, bwd
    # ---- K1-fused: to_q -> masked attention -> to_out in one cluster kernel ----------------------
    x_probe = torch.empty(B, T, D, device=dev, dtype=dtype)
    kv_probe = torch.empty(B, Ti * n, 2 * inner, device=dev, dtype=dtype)
    if ops.xattn_block_supported(x_probe, kv_probe, heads=H, n_latents=n):
        fb = es * (2 * B * T * D + 2 * B * T * inner + 2 * B * Ti * n * inner + 2 * D * inner) + 4 * B * T * H
        ff = 2.0 * B * T * D * inner * 2 + xf
        K = _k(fb, cap=24)
        xs = [torch.randn(B, T, D, device=dev, dtype=dtype) for _ in range(K)]
        kvs = [torch.randn(B, Ti * n, 2 * inner, device=dev, dtype=dtype) for _ in range(K)]
        wqs = [(torch.randn(inner, D, device=dev) * D ** -0.5).to(dtype) for _ in range(K)]
        wos = [(torch.randn(D, inner, device=dev) * inner ** -0.5).to(dtype) for _ in range(K)]
        with torch.no_grad():
            rec("xattn_block_fwd", _time_graph([
                lambda x=x, kv=kv, wq=wq, wo=wo: ops.xattn_block(x, wq, kv, tt, wo, heads=H, n_latents=n, scale=dh ** -0.5)
                for x, kv, wq, wo in zip(xs, kvs, wqs, wos)]), fb, ff, K)
            rec("xattn_three_launches", _time_graph([
                lambda x=x, kv=kv, wq=wq, wo=wo: torch.nn.functional.linear(ops.masked_cross_attention(
                    torch.nn.functional.linear(x, wq), kv, tt, heads=H, n_latents=n, scale=dh ** -0.5), wo)
                for x, kv, wq, wo in zip(xs, kvs, wqs, wos)]), fb, ff, K)
            if eager:
                rec("eager_xattn_block_fwd", _time_graph([
                    lambda x=x, kv=kv, wq=wq, wo=wo: torch.nn.functional.linear(_eager_xattn(
                        torch.nn.functional.linear(x, wq), kv, tt, H, n, dh ** -0.5), wo)
                    for x, kv, wq, wo in zip(xs, kvs, wqs, wos)]), fb, ff, K)
        del xs, kvs, wqs, wos


def _run_lm_attn(cfg, wl, dtype, eager, rec):
    """K4: causal self-attention of one GPT-NeoX layer on the packed, rotated qkv projection
    (B, T, H, 3, dh); eager column = what the reference runs (HF sdpa attention -> cuDNN flash)."""
    dev, es, B, T = "cuda", 2, wl.B, wl.T
    H, dh = cfg.lm_heads, cfg.lm_hidden // cfg.lm_heads
    probe = torch.empty(B, H, T, dh, device=dev, dtype=dtype)
    if dtype != torch.bfloat16 or not ops.lm_attention_supported(probe):
        return
    byts = es * 4 * B * T * H * dh + 4 * B * H * T          # q, k, v, o + lse
    fl = 4.0 * B * H * dh * T * (T + 1) / 2                 # causal half
    K = _k(byts, cap=16)
    packs = [torch.randn(B, T, H * 3 * dh, device=dev, dtype=dtype) for _ in range(K)]
    gos = [torch.randn(B, T, H * dh, device=dev, dtype=dtype) for _ in range(K)]

    def views(p):   # leaf q/k/v VIEWS of the packed projection: the backward numbers carry no un-packing glue
        return tuple(p.view(B, T, H, 3, dh)[:, :, :, i].transpose(1, 2).requires_grad_() for i in range(3))

    # right-padded batch (the workloads' batches are: SURVEY.md §8 synthetic inputs), one full-length sample
    lens = torch.linspace(T // 2, T, B).long()
    km = (torch.arange(T)[None, :] < lens[:, None]).to(dev)
    bits = ops.key_bits(km)
    qkvs = [views(p) for p in packs]
    fwd = [lambda x=x: ops.lm_attention(*x, bits, scale=dh ** -0.5) for x in qkvs]
    rec("lm_attn_fwd", _time_graph(fwd), byts, fl, K)
    os_ = [f() for f in fwd]
    bwd = [lambda o=o, x=x, g=g: torch.autograd.grad(o, x, g, retain_graph=True) for o, x, g in zip(os_, qkvs, gos)]
    rec("lm_attn_bwd", _time_graph(bwd), 2.5 * byts, 2.5 * fl, K)
    if eager:
        # what the reference runs: HF sdpa attention with the 4-D boolean mask HF builds from a padded
        # 2-D attention_mask (causal & key padding); and the best case (no padding: is_causal=True)
        F = torch.nn.functional
        mask4 = torch.ones(T, T, dtype=torch.bool, device=dev).tril()[None, None] & km[:, None, None, :]
        for tag, kw in (("", dict(attn_mask=mask4)), ("_causal_only", dict(is_causal=True))):
            efw = [lambda x=x: F.scaled_dot_product_attention(*x, scale=dh ** -0.5, **kw)
                   .transpose(1, 2).reshape(B, T, H * dh) for x in qkvs]
            rec(f"eager_lm_attn_fwd{tag}", _time_graph(efw), byts, fl, K)
            eos = [f() for f in efw]
            rec(f"eager_lm_attn_bwd{tag}",
                _time_graph([lambda o=o, x=x, g=g: torch.autograd.grad(o, x, g, retain_graph=True)
                             for o, x, g in zip(eos, qkvs, gos)]), 2.5 * byts, 2.5 * fl, K)
            del efw, eos


def _run_vit_perceiver(cfg, wl, dtype, eager, rec):
    dev, es, B, T, Ti, n, H, dh, D, V = _dims(cfg, wl, dtype)
    inner = H * dh
    # ---- K3 ViT self-attention ----------------------------------------------------------------
    N, L, Hv = B * Ti, cfg.n_patches + 1, cfg.vis_heads
    vb = es * 4 * N * L * Hv * 64 + 4 * N * L * Hv
    vf = 4.0 * N * Hv * L * L * 64
    K = _k(vb)
    qkvs = [torch.randn(N, L, 3 * Hv * 64, device=dev, dtype=dtype) for _ in range(K)]
    rec("vit_attn_fwd", _time_graph([lambda x=x: ops.attention(x[..., :Hv * 64], x[..., Hv * 64:], heads=Hv, scale=0.125)
                                     for x in qkvs]), vb, vf, K)
    if eager:
        def sdpa(x):
            q_, k_, v_ = (t.view(N, L, Hv, 64).transpose(1, 2) for t in x.chunk(3, dim=-1))
            return torch.nn.functional.scaled_dot_product_attention(q_, k_, v_, scale=0.125)
        rec("eager_vit_attn_fwd_sdpa", _time_graph([lambda x=x: sdpa(x) for x in qkvs]), vb, vf, K)
    del qkvs
    # ---- K2 perceiver attention -----------------------------------------------------------------
    Lk = cfg.n_patches + n
    pb = es * (2 * N * n * inner + 2 * N * Lk * inner) + 4 * N * n * H
    pf = 4.0 * N * H * n * Lk * dh
    K = _k(pb)
    pqs = [torch.randn(N, n, inner, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    pkvs = [torch.randn(N, Lk, 2 * inner, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    pfw = [lambda q=q, kv=kv: ops.attention(q, kv, heads=H, scale=dh ** -0.5) for q, kv in zip(pqs, pkvs)]
    rec("perceiver_attn_fwd", _time_graph(pfw), pb, pf, K)
    pos = [f() for f in pfw]
    pgs = [torch.randn_like(o) for o in pos]
    rec("perceiver_attn_bwd", _time_graph([lambda o=o, q=q, kv=kv, g=g: torch.autograd.grad(o, (q, kv), g, retain_graph=True)
                                            for o, q, kv, g in zip(pos, pqs, pkvs, pgs)]), 2.5 * pb, 2.5 * pf, K)
    if eager:
        epf = [lambda q=q, kv=kv: _eager_attn(q, kv, H, dh ** -0.5) for q, kv in zip(pqs, pkvs)]
        rec("eager_perceiver_attn_fwd", _time_graph(epf), pb, pf, K)
        epo = [f() for f in epf]
        rec("eager_perceiver_attn_bwd", _time_graph([lambda o=o, q=q, kv=kv, g=g: torch.autograd.grad(o, (q, kv), g, retain_graph=True)
                                                      for o, q, kv, g in zip(epo, pqs, pkvs, pgs)]), 2.5 * pb, 2.5 * pf, K)
        del epf, epo
    del pqs, pkvs, pos, pgs, pfw


def _run_k5_gelu(cfg, wl, dtype, eager, rec):
    dev, es, B, T, Ti, n, H, dh, D, V = _dims(cfg, wl, dtype)
    # ---- K5 gate + residual + LN ------------------------------------------------------------------
    rows = B * T
    gb = 4 * rows * D * es
    K = _k(gb, cap=16)
    xs = [torch.randn(rows, D, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    brs = [torch.randn(rows, D, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    gate = torch.full((1,), 0.5, device=dev, dtype=dtype, requires_grad=True)
    gam = torch.ones(D, device=dev, dtype=dtype, requires_grad=True)
    bet = torch.zeros(D, device=dev, dtype=dtype, requires_grad=True)
    gfw = [lambda x=x, b=b: ops.gate_residual_ln(b, x, gate, gam, bet) for x, b in zip(xs, brs)]
    rec("gate_residual_ln_fwd", _time_graph(gfw), gb, 0.0, K)
    outs = [f() for f in gfw]
    g1s = [torch.randn(rows, D, device=dev, dtype=dtype) for _ in range(K)]
    rec("gate_residual_ln_bwd", _time_graph([lambda o=o, x=x, b=b, g=g: torch.autograd.grad(
        o, (b, x, gate, gam, bet), (g, g), retain_graph=True) for o, x, b, g in zip(outs, xs, brs, g1s)]),
        6 * rows * D * es, 0.0, K)
    if eager:
        def eg(b, x):
            xo = b * gate.tanh() + x
            return xo, torch.nn.functional.layer_norm(xo, (D,), gam, bet)
        egf = [lambda x=x, b=b: eg(b, x) for x, b in zip(xs, brs)]
        rec("eager_gate_residual_ln_fwd", _time_graph(egf), gb, 0.0, K)
        eouts = [f() for f in egf]
        rec("eager_gate_residual_ln_bwd", _time_graph([lambda o=o, x=x, b=b, g=g: torch.autograd.grad(
            o, (b, x, gate, gam, bet), (g, g), retain_graph=True) for o, x, b, g in zip(eouts, xs, brs, g1s)]),
            6 * rows * D * es, 0.0, K)
        del egf, eouts
    del outs, gfw
    # the LM towers' variant: ungated residual + frozen LayerNorm (no column sums, d_branch == d_x)
    gam_f, bet_f = gam.detach(), bet.detach()
    outs = [ops.gate_residual_ln(b, x, None, gam_f, bet_f) for x, b in zip(xs, brs)]
    rec("residual_ln_bwd_frozen", _time_graph([lambda o=o, x=x, b=b, g=g: torch.autograd.grad(
        o, (b, x), (g, g), retain_graph=True) for o, x, b, g in zip(outs, xs, brs, g1s)]),
        4 * rows * D * es, 0.0, K)
    del xs, brs, outs, g1s
    # ---- exact GELU of the FeedForward blocks (rows x 4D) ---------------------------------------
    F4 = cfg.ff_mult * D
    K = _k(2 * rows * F4 * es, cap=8)
    hs = [torch.randn(rows, F4, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
    rec("gelu_fwd", _time_graph([lambda h=h: ops.gelu(h) for h in hs]), 2 * rows * F4 * es, 0.0, K)
    ys = [ops.gelu(h) for h in hs]
    gys = [torch.randn(rows, F4, device=dev, dtype=dtype) for _ in range(K)]
    rec("gelu_bwd", _time_graph([lambda y=y, h=h, g=g: torch.autograd.grad(y, h, g, retain_graph=True)
                                 for y, h, g in zip(ys, hs, gys)]), 3 * rows * F4 * es, 0.0, K)
    # the stock kernels at the same shape, for the record (not on the product path)
    rec("eager_gelu_fwd", _time_graph([lambda h=h: torch.nn.functional.gelu(h) for h in hs]),
        2 * rows * F4 * es, 0.0, K)
    ts = [torch.nn.functional.gelu(h) for h in hs]
    rec("eager_gelu_bwd", _time_graph([lambda y=y, h=h, g=g: torch.autograd.grad(y, h, g, retain_graph=True)
                                       for y, h, g in zip(ts, hs, gys)]), 3 * rows * F4 * es, 0.0, K)
    del hs, ys, ts, gys


def _run_focal(cfg, wl, dtype, eager, rec, out, n_valid_rows):
    dev, es, B, T, Ti, n, H, dh, D, V = _dims(cfg, wl, dtype)
    # ---- K6 focal CE -----------------------------------------------------------------------------
    nv = n_valid_rows if n_valid_rows else max(1, B * (Ti + 2))
    z = [torch.randn(B, T, V, device=dev, dtype=dtype, requires_grad=True) for _ in range(2)]
    y = torch.full((B * T,), -100, device=dev, dtype=torch.int64)
    idx = torch.randperm(B * T - B, device=dev)[:nv] + 1
    y[idx] = torch.randint(0, V, (nv,), device=dev)
    y = y.view(B, T)
    y[:, 0] = -100
    nv_eff = int((y[:, 1:] != -100).sum())
    w = torch.ones(B, device=dev)
    ffw = [lambda zz=zz: ops.focal_ce(zz, y, w) for zz in z]
    rec("focal_ce_fwd", _time_graph(ffw), nv_eff * V * es, 0.0, 2)
    ls = [f() for f in ffw]
    rec("focal_ce_bwd", _time_graph([lambda l=l, zz=zz: torch.autograd.grad(l, zz, retain_graph=True)
                                     for l, zz in zip(ls, z)]), nv_eff * V * es + B * T * V * es, 0.0, 2)
    out["focal_ce_fwd"]["n_valid_rows"] = nv_eff
    if eager:
        w2 = torch.full((B,), 2.0, device=dev)
        efw = [lambda zz=zz: _eager_focal(zz, y, w2, 2.0) for zz in z]
        rec("eager_focal_ce_fwd", _time_graph(efw), nv_eff * V * es, 0.0, 2)
        els = [f() for f in efw]
        rec("eager_focal_ce_bwd", _time_graph([lambda l=l, zz=zz: torch.autograd.grad(l, zz, retain_graph=True)
                                               for l, zz in zip(els, z)]), nv_eff * V * es + B * T * V * es, 0.0, 2)
        del efw, els
    del z, ls, ffw
    # ---- head + loss fusion: the same loss over the gathered valid rows only ---------------------
    Rcap = max(64, (nv_eff + 63) // 64 * 64)
    Vp = (V + 127) // 128 * 128
    K = _k(Rcap * Vp * es, cap=24)
    zr = [torch.randn(Rcap, Vp, device=dev, dtype=dtype)[:, :V].requires_grad_(True) for _ in range(K)]
    tg = torch.full((Rcap,), -100, device=dev, dtype=torch.int64)
    tg[:nv_eff] = torch.randint(0, V, (nv_eff,), device=dev)
    rw = torch.ones(Rcap, device=dev)
    rfw = [lambda zz=zz: ops.focal_ce_rows(zz, tg, rw, None) for zz in zr]
    rec("focal_ce_rows_fwd", _time_graph(rfw), nv_eff * V * es, 0.0, K)
    rls = [f() for f in rfw]
    rec("focal_ce_rows_bwd", _time_graph([lambda l=l, zz=zz: torch.autograd.grad(l, zz, retain_graph=True)
                                          for l, zz in zip(rls, zr)]), nv_eff * V * es + Rcap * V * es, 0.0, K)
    del zr, rls, rfw


def _run_misc(cfg, wl, dtype, rec):
    dev, es, B, T, Ti, n, H, dh, D, V = _dims(cfg, wl, dtype)
    # ---- rotary on the packed qkv projection (GPT-NeoX, 32 heads x 80, rotary_pct 1.0) -----------
    Hl, dl = cfg.lm_heads, cfg.lm_hidden // cfg.lm_heads
    rot = int(dl * cfg.rotary_pct)
    if rot % 16 == 0 and dl % 8 == 0:
        rb = 2 * B * T * 3 * Hl * dl * es
        K = _k(rb, cap=16)
        qkvs = [torch.randn(B, T, 3 * Hl * dl, device=dev, dtype=dtype, requires_grad=True) for _ in range(K)]
        cos = torch.randn(1, T, rot, device=dev, dtype=dtype)
        sin = torch.randn(1, T, rot, device=dev, dtype=dtype)
        rf = [lambda x=x: ops.rotary_qkv(x, cos, sin, heads=Hl, head_dim=dl, rotary_dim=rot) for x in qkvs]
        rec("rotary_qkv_fwd", _time_graph(rf), rb, 0.0, K)
        ro = [f() for f in rf]
        gq = [tuple(torch.randn_like(t) for t in o) for o in ro]
        rec("rotary_qkv_bwd", _time_graph([lambda o=o, x=x, g=g: torch.autograd.grad(o, x, g, retain_graph=True)
                                           for o, x, g in zip(ro, qkvs, gq)]), rb, 0.0, K)
        del qkvs, ro, gq, rf
    # ---- clip + AdamW over a 64 M-parameter slice (28 B/param; the step runs it over 1.15 B) -----
    npar = 64 << 20
    pw = torch.randn(npar, device=dev, dtype=dtype)
    pg = torch.randn(npar, device=dev, dtype=dtype)
    ma, m1, m2 = pw.float(), torch.zeros(npar, device=dev), torch.zeros(npar, device=dev)
    hyper = torch.tensor(ops.adamw_hyper(2e-4, 0.9, 0.999, 10), device=dev, dtype=torch.float32)
    gn = torch.ones(1, device=dev)
    rec("adamw_clip_step", _time_graph([lambda: ops.adamw_step_(ma, pw, pg, m1, m2, hyper=hyper, beta1=0.9,
                                                                 beta2=0.999, eps=1e-8, weight_decay=0.1,
                                                                 gnorm_sq=gn, max_norm=1.0)] * 3),
        npar * (2 * es + 6 * 4), 0.0, 3)
    rec("adamw_clip_step_background", _time_graph([lambda: ops.adamw_step_(ma, pw, pg, m1, m2, hyper=hyper, beta1=0.9,
                                                                            beta2=0.999, eps=1e-8, weight_decay=0.1,
                                                                            gnorm_sq=gn, max_norm=1.0, background=True)] * 3),
        npar * (2 * es + 6 * 4), 0.0, 3)
    rec("grad_sumsq", _time_graph([lambda: ops.sumsq_(pg, gn)] * 3), npar * es, 0.0, 3)
    del pw, pg, ma, m1, m2


def _run_decode(cfg, dtype, eager, rec, beams=5, cursor=640, t_max=768, ti=5):
    """The decode step's own kernels at configs[3] shapes (batch 1 x 5 beams, prompt 512 + 128 generated
    tokens in the caches), each beside what the library path launches for the same work."""
    F = torch.nn.functional
    dev = "cuda"
    D, Hl = cfg.lm_hidden, cfg.lm_heads
    dl = D // Hl
    rot = int(dl * cfg.rotary_pct)
    M = beams
    with torch.no_grad():
        # ---- projections of one token: M rows against a cold weight matrix ------------------------
        inner = cfg.xattn_heads * cfg.xattn_dim_head
        for name, N, K in (("qkv", 3 * D, D), ("dense", D, D), ("h_to_4h", cfg.lm_ffn, D), ("4h_to_h", D, cfg.lm_ffn),
                           ("to_q", inner, D), ("to_out", D, inner)):
            nb = _k(N * K * 2, cap=64)
            ws = [torch.randn(N, K, device=dev, dtype=dtype) * K ** -0.5 for _ in range(nb)]
            b = torch.randn(N, device=dev, dtype=dtype)
            x = torch.randn(M, 1, K, device=dev, dtype=dtype)
            assert ops.small_m_eligible(x, ws[0])
            rec("decode_linear_" + name, _time_graph([lambda w=w: ops.linear_rows(x, w, b) for w in ws]), N * K * 2, 0.0, nb)
            if eager:
                rec("eager_decode_linear_" + name, _time_graph([lambda w=w: F.linear(x, w, b) for w in ws]), N * K * 2, 0.0, nb)
            del ws
        # ---- the LM self-attention of one token: rotary + cache write + attention over `cursor` keys --
        kvb = 2 * M * Hl * cursor * dl * 2
        nb = _k(kvb, cap=12)
        kc = [torch.randn(M, Hl, t_max, dl, device=dev, dtype=dtype) for _ in range(nb)]
        vc = [torch.randn(M, Hl, t_max, dl, device=dev, dtype=dtype) for _ in range(nb)]
        qkv = torch.randn(M, 1, 3 * D, device=dev, dtype=dtype)
        cos, sin = torch.randn(M, 1, rot, device=dev, dtype=dtype), torch.randn(M, 1, rot, device=dev, dtype=dtype)
        indir = torch.randint(0, M, (M, t_max), device=dev, dtype=torch.int32)
        cur = torch.tensor([cursor], device=dev, dtype=torch.int64)
        mask = torch.zeros(M, t_max, device=dev, dtype=dtype)
        mask[:, cursor + 1:] = float("-inf")
        rec("lm_decode_attn", _time_graph([lambda k=k, v=v: ops.lm_decode_attention(
            qkv, cos, sin, k, v, indir, mask, cur, heads=Hl, head_dim=dl, rotary_dim=rot, scale=dl ** -0.5)
            for k, v in zip(kc, vc)]), kvb, 0.0, nb)
        if eager:
            src = torch.randint(0, M, (M,), device=dev)
            m4 = mask.view(M, 1, 1, t_max)

            def lib_step(k, v):      # what HF's generate does per layer and token on this GPU
                q, kn, vn = ops.rotary_qkv(qkv, cos, sin, heads=Hl, head_dim=dl, rotary_dim=rot)
                k.index_copy_(2, cur, kn)
                v.index_copy_(2, cur, vn)
                a = F.scaled_dot_product_attention(q, k, v, attn_mask=m4, scale=dl ** -0.5)
                k.copy_(k.index_select(0, src))      # reorder_cache after the beam step
                v.copy_(v.index_select(0, src))
                return a

            rec("eager_lm_decode_attn", _time_graph([lambda k=k, v=v: lib_step(k, v) for k, v in zip(kc, vc)]),
                kvb, 0.0, nb)
        del kc, vc
        # ---- single-token masked cross-attention against the cached to_kv(media) ------------------
        n, Hx, dx = cfg.n_latents, cfg.xattn_heads, cfg.xattn_dim_head
        q1 = torch.randn(M, 1, Hx * dx, device=dev, dtype=dtype)
        kvs = [torch.randn(M, ti * n, 2 * Hx * dx, device=dev, dtype=dtype) for _ in range(16)]
        nm = torch.full((M,), ti, device=dev, dtype=torch.int32)
        xb = 2 * M * n * Hx * dx * 2
        rec("xattn_decode", _time_graph([lambda kv=kv: ops.xattn_decode(q1, kv, nm, heads=Hx, n_latents=n, scale=dx ** -0.5)
                                         for kv in kvs]), xb, 0.0, len(kvs))
        if eager:
            tt = nm[:, None]
            rec("eager_xattn_decode", _time_graph([lambda kv=kv: _eager_xattn(q1, kv, tt, Hx, n, dx ** -0.5) for kv in kvs]),
                xb, 0.0, len(kvs))
