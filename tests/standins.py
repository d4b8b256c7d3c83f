"""TEST DOUBLES for the CUDA ops — test infrastructure, never imported by unimp_b200/.

`install()` replaces the functions of `unimp_b200.ops` with dense PyTorch statements of the same
contracts (the formulas the GPU parity tests check the kernels against) and makes the CUDA-graph
capture a plain closure, so that HOST logic which normally needs a GPU — the CUDA-graph decoder's
prefill / static K-V caches / masks / positions / beam re-ordering — can be exercised on a GPU-less
box.  It also makes `Tensor.is_cuda` report True, so it must only ever be used in a throw-away
subprocess (`python tests/standins.py decode`), never inside the pytest process.  The product has
no CPU path: without `install()` every op raises on CPU tensors (tests/test_capi_symbols.py).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _dense_attention(q, kv, tt, heads, n, scale):
    B, Lq, inner = q.shape
    Lk, dh = kv.shape[1], inner // heads
    k, v = kv[..., :inner], kv[..., inner:]
    qh = q.reshape(B, Lq, heads, dh).transpose(1, 2) * scale
    kh = k.reshape(B, Lk, heads, dh).transpose(1, 2)
    vh = v.reshape(B, Lk, heads, dh).transpose(1, 2)
    sim = qh @ kh.transpose(-1, -2)
    if tt is not None:
        media_time = (torch.arange(Lk // n) + 1).repeat_interleave(n)
        mask = tt[:, None, :, None] == media_time[None, None, None, :]
        sim = sim.masked_fill(~mask, -torch.finfo(sim.dtype).max)
    attn = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    if tt is not None:
        attn = attn.masked_fill((tt == 0)[:, None, :, None], 0.0)
    return (attn @ vh).transpose(1, 2).reshape(B, Lq, inner)


def install():
    import unimp_b200.decode as decode
    import unimp_b200.ops as ops

    def text_time(lang_x, media_token_id, *, use_cached=False, T_out=None):
        loc = lang_x == media_token_id
        if use_cached:
            return loc.sum(-1, keepdim=True).expand(lang_x.shape[0], T_out).to(torch.int32).contiguous()
        return loc.cumsum(-1).to(torch.int32)

    def gate_residual_ln(branch, x, gate, gamma, beta, eps=1e-5):
        xo = branch * (gate.tanh() if gate is not None else 1.0) + x
        return xo, F.layer_norm(xo, (x.shape[-1],), gamma, beta, eps)

    def rotary_qkv(qkv, cos, sin, *, heads, head_dim, rotary_dim):
        from transformers.models.gpt_neox.modeling_gpt_neox import apply_rotary_pos_emb

        B, T, _ = qkv.shape
        v5 = qkv.view(B, T, heads, 3, head_dim)
        q, k, v = (v5[:, :, :, i].transpose(1, 2) for i in range(3))
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        return q, k, v

    def lm_decode_attention(qkv, cos, sin, k_cache, v_cache, indir, add_mask, cursor, *, heads, head_dim,
                            rotary_dim, scale):
        """The contract of `unimp_lm_decode_attn`, densely: rotary, cache write at the cursor, attention
        over positions 0..cursor with position t of beam b read from cache row indir[b, t]."""
        from transformers.models.gpt_neox.modeling_gpt_neox import apply_rotary_pos_emb

        B, Tmax = qkv.shape[0], k_cache.shape[2]
        v5 = qkv.view(B, 1, heads, 3, head_dim)
        q, k, v = (v5[:, :, :, i].transpose(1, 2) for i in range(3))          # (B,H,1,dh)
        q, k = apply_rotary_pos_emb(q, k, cos.view(B, 1, rotary_dim), sin.view(B, 1, rotary_dim))
        cur = int(cursor)
        k_cache[:, :, cur], v_cache[:, :, cur] = k[:, :, 0], v[:, :, 0]
        indir[:, cur] = torch.arange(B, dtype=torch.int32)
        rows, t = indir.long(), torch.arange(Tmax)
        kg, vg = k_cache[rows, :, t[None, :]], v_cache[rows, :, t[None, :]]   # (B,Tmax,H,dh)
        s = torch.einsum("bhd,bthd->bht", q[:, :, 0], kg) * scale + add_mask.view(B, 1, Tmax)
        s[:, :, cur + 1:] = float("-inf")
        return torch.einsum("bht,bthd->bhd", s.softmax(-1), vg).reshape(B, 1, heads * head_dim)

    def linear_rows(x, w, b=None, act_gelu=False):
        y = F.linear(x, w, b)
        return F.gelu(y) if act_gelu else y

    ops.lm_decode_attention = lm_decode_attention
    ops.linear_rows = linear_rows
    ops.small_m_eligible = lambda x, w: False
    ops.BEAM_TOPK = False                    # candidate selection stays on torch's log_softmax + topk here
    ops.text_time = text_time
    ops.masked_cross_attention = lambda q, kv, tt, *, heads, n_latents, scale, force_simt=False: \
        _dense_attention(q, kv, tt.long(), heads, n_latents, scale)
    ops.attention = lambda q, kv, *, heads, scale, force_simt=False: _dense_attention(q, kv, None, heads, 1, scale)
    ops.xattn_decode = lambda q, kv, n_media, *, heads, n_latents, scale: \
        _dense_attention(q, kv, n_media.long()[:, None], heads, n_latents, scale)
    ops.gate_residual_ln = gate_residual_ln
    ops.gate_residual = lambda b, x, g: b * (g.tanh() if g is not None else 1.0) + x
    ops.layer_norm = lambda x, g, b, eps=1e-5: F.layer_norm(x, (x.shape[-1],), g, b, eps)
    ops.rotary_qkv = rotary_qkv
    ops.quick_gelu_ = lambda x: x.mul_(torch.sigmoid(1.702 * x))
    ops.gelu = F.gelu
    ops.linear_acc = lambda x, w: F.linear(x, w)
    ops.embedding_acc = lambda ids, w: F.embedding(ids, w)
    ops.key_bits = lambda mask: mask            # the double below takes the 2-D mask itself
    ops.lm_attention_supported = lambda q: False  # K4 stays on SDPA here (fp32 tiny model anyway)
    ops.lm_attention_supported_shape = lambda x, T, H, dh: False

    class _Closure:   # a "graph" that simply re-runs the captured step
        def __init__(self, fn):
            self.fn = fn

        def replay(self):
            self.fn()

    decode.GraphedDecoder._capture = staticmethod(lambda fn: _Closure(fn))
    torch.Tensor.is_cuda = property(lambda self: True)


def decode_check():
    """GraphedDecoder (host logic on the doubles) against Flamingo.generate (HF generate on the same
    doubles): identical token matrices, full and ragged prompts, greedy and beams."""
    install()
    from unimp_b200 import tiny_config
    from unimp_b200.config import Workload
    from unimp_b200.decode import GraphedDecoder
    from unimp_b200.factory import build_flamingo
    from unimp_b200.synth import make_batch

    cfg = tiny_config()
    model = build_flamingo(cfg, dtype=torch.float32, device="cpu", gate=0.5).eval()
    b = make_batch(cfg, Workload("dec", B=2, Ti=2, T=40), seed=3)
    L = int(b["attention_masks"].sum(-1).min()) - 2
    ids, vis = b["input_ids"][:, :L], b["patch_images"].unsqueeze(2)
    eos, pad = cfg.tokens.endofchunk, cfg.tokens.pad
    ok = True
    cases = [(ids, torch.ones_like(ids), nb, nrs, es, new)
             for nb, nrs, es, new in [(1, 1, False, 9), (3, 2, True, 9), (5, 5, True, 12), (4, 1, False, 7),
                                      (2, 2, True, 1), (3, 1, True, 2)]]
    for side in ("right", "left"):
        ids2, m2 = ids.clone(), torch.ones_like(ids)
        sl = slice(-5, None) if side == "right" else slice(0, 5)
        ids2[1, sl], m2[1, sl] = pad, 0
        cases += [(ids2, m2, 1, 1, True, 8), (ids2, m2, 3, 2, True, 8)]
    for x, m, nb, nrs, es, new in cases:
        kw = dict(max_new_tokens=new, eos_token_id=eos, pad_token_id=pad)
        if nb > 1:
            kw.update(num_beams=nb, num_return_sequences=nrs, early_stopping=es)
        want = model.generate(vision_x=vis, lang_x=x, attention_mask=m, do_sample=False, **kw)
        got = GraphedDecoder(model, sync_every=4).generate(vis, x, m, num_beams=nb, max_new_tokens=new,
                                                          eos_token_id=eos, pad_token_id=pad,
                                                          num_return_sequences=nrs, early_stopping=es)
        same = got.shape == want.shape and torch.equal(got, want)
        print(f"beams={nb} returned={nrs} early_stopping={es} new={new} padded={int((m == 0).any())}: "
              f"{'equal' if same else 'DIFFERENT'}")
        ok &= same
    print("DECODE CHECK " + ("ALL EQUAL" if ok else "FAILED"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(decode_check() if sys.argv[1:] == ["decode"] else 2)
