"""PerceiverResampler / MaskedCrossAttention / GatedCrossAttentionBlock on the sm_100a kernels.

Same class names, constructor arguments, parameter names (=> state-dict keys) and forward
semantics as `open_flamingo/src/helpers.py` v2.0.1 (the reference imports them at
`UniMP/mmrec.py:20-22` and names their parameters at `UniMP/mmrec.py:612-619`).  What differs is
how forward runs: the attention cores, the gate+residual+LayerNorm epilogues and every
LayerNorm are the hand-written kernels behind `unimp_b200.ops`; only the dense projections
stay on cuBLAS (`F.linear`).  No autocast: activations live in the parameters' dtype.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def exists(v):
    return v is not None


def FeedForward(dim: int, mult: int = 4) -> nn.Sequential:
    inner = int(dim * mult)
    return nn.Sequential(
        nn.LayerNorm(dim),
        nn.Linear(dim, inner, bias=False),
        nn.GELU(),
        nn.Linear(inner, dim, bias=False),
    )


def _ff_tail(ff: nn.Sequential, ln_out: torch.Tensor) -> torch.Tensor:
    """Linear -> GELU -> Linear of a FeedForward whose LayerNorm was already applied (fused)."""
    if ops.small_m_eligible(ln_out, ff[1].weight):     # a decode step: weight-streaming kernels
        return ops.linear_rows(ops.linear_rows(ln_out, ff[1].weight, None, act_gelu=True), ff[3].weight)
    h = ops.linear_acc(ln_out, ff[1].weight)
    h = ops.gelu(h)
    return ops.linear_acc(h, ff[3].weight)


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner = dim_head * heads
        self.norm_media = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, latents, latents_ln=None):
        """x (b,T,n1,D), latents (b,T,n2,D).  `latents_ln`: norm_latents(latents) if the caller
        already produced it in a fused epilogue."""
        b, T, n1, D = x.shape
        n2 = latents.shape[2]
        x_ln = ops.layer_norm(x, self.norm_media.weight, self.norm_media.bias, self.norm_media.eps)
        if latents_ln is None:
            latents_ln = ops.layer_norm(latents, self.norm_latents.weight, self.norm_latents.bias,
                                        self.norm_latents.eps)
        q = ops.linear_acc(latents_ln, self.to_q.weight).view(b * T, n2, -1)
        kv_in = torch.cat((x_ln, latents_ln), dim=-2)
        kv = ops.linear_acc(kv_in, self.to_kv.weight).view(b * T, n1 + n2, -1)
        out = ops.attention(q, kv, heads=self.heads, scale=self.scale)      # K2
        return ops.linear_acc(out, self.to_out.weight).view(b, T, n2, D)


class PerceiverResampler(nn.Module):
    def __init__(self, *, dim, depth=6, dim_head=64, heads=8, num_latents=64, max_num_media=None,
                 max_num_frames=None, ff_mult=4):
        super().__init__()
        if exists(max_num_media) or exists(max_num_frames):
            raise ValueError("frame / media-time embeddings are not used by Flamingo.__init__ "
                             "(it passes only `dim`) and are not implemented")
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.frame_embs = None
        self.media_time_embs = None
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads),
                FeedForward(dim=dim, mult=ff_mult),
            ]))
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        """x (b, T, F, v, D) -> (b, T, n, D)."""
        b, T, Fr, v, D = x.shape
        x = x.reshape(b, T, Fr * v, D).contiguous()  # once (the ViT output is a CLS-dropping slice)
        latents = self.latents.to(x.dtype).expand(b, T, -1, -1).contiguous()
        latents_ln = None
        n_layers = len(self.layers)
        for li, (attn, ff) in enumerate(self.layers):
            a = attn(x, latents, latents_ln)
            # latents = a + latents, fused with the FF's LayerNorm (K5, no gate)
            latents, h = ops.gate_residual_ln(a, latents, None, ff[0].weight, ff[0].bias, ff[0].eps)
            h = _ff_tail(ff, h)
            # latents = ff + latents, fused with the LayerNorm that reads it next
            nxt = self.layers[li + 1][0].norm_latents if li + 1 < n_layers else self.norm
            latents, latents_ln = ops.gate_residual_ln(h, latents, None, nxt.weight, nxt.bias,
                                                       nxt.eps)
        return latents_ln  # == self.norm(latents)


class MaskedCrossAttention(nn.Module):
    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, only_attend_immediate_media=True):
        super().__init__()
        if not only_attend_immediate_media:
            raise ValueError("only_attend_immediate_media=False is not on the UniMP path "
                             "(GatedCrossAttentionBlock's default is True)")
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        inner = dim_head * heads
        self.norm = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim_visual, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        self.only_attend_immediate_media = True
        self._kv_cache = None  # decode-time cache of to_kv(media) (new)
        # K1-fused (one cluster kernel for to_q -> attention -> to_out) where the shape allows (bf16,
        # 8 x 64 heads, 64 latents, D <= 2560) AND its 8-CTA clusters run as one wave: 12 clusters are
        # co-resident on a B200, i.e. B * ceil(T / 128) <= 12 row tiles (configs[1]: 12).  Beyond that
        # the clusters queue in waves and cuBLAS + core + cuBLAS is faster (configs[2], 48 tiles: 103 vs
        # 46 us in isolation, -0.7 % samples/s in the step).  UNIMP_XATTN_FUSED=0 never fuses, =2 always.
        self.fused = int(os.environ.get("UNIMP_XATTN_FUSED", "1"))

    FUSED_MAX_TILES = 12   # co-resident 8-CTA clusters per B200

    def project_media(self, media):
        """to_kv over (B, Ti*n, Dv) -> packed (B, Ti*n, 2*inner)."""
        B, Ti, n, Dv = media.shape
        return ops.linear_acc(media.reshape(B, Ti * n, Dv), self.to_kv.weight)

    def cached_media_kv(self, media):
        """to_kv(media), computed once per (media tensor, weight version) — decode-time cache."""
        key = (id(media), media._version, self.to_kv.weight._version, self.to_kv.weight.data_ptr(),
               ops.weights_epoch())   # FlatAdamW updates weights through raw pointers: no _version bump
        if self._kv_cache is None or self._kv_cache[0] != key:
            self._kv_cache = (key, self.project_media(media), media)  # `media` kept alive: id stays unique
        return self._kv_cache[1]

    def forward(self, x, media, media_locations=None, use_cached_media=False, text_time=None,
                x_ln=None):
        """x (B,T,D); media (B,Ti,n,Dv); `text_time` int32 (B,T) may be passed precomputed
        (FlamingoLMMixin does, once per forward); otherwise derived from media_locations.
        `x_ln`: self.norm(x) if the producer of x already emitted it from a fused epilogue."""
        B, T, D = x.shape
        _, Ti, n = media.shape[:3]
        if text_time is None:
            if not exists(media_locations):
                raise ValueError("media_locations (or text_time) is required: the UniMP path "
                                 "always conditions media locations")
            if not use_cached_media:
                assert media_locations.shape[1] == T, (
                    f"media_location.shape is {media_locations.shape} but x.shape is {x.shape}")
            text_time = ops.text_time(media_locations.to(torch.int64), 1,
                                      use_cached=use_cached_media, T_out=T)
        if x_ln is None:
            x_ln = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        cached = use_cached_media and not torch.is_grad_enabled()
        fuse = self.fused == 2 or (self.fused == 1 and B * ((T + 127) // 128) <= self.FUSED_MAX_TILES)
        if fuse and not (cached and T == 1):
            kv = self.cached_media_kv(media) if cached else self.project_media(media)
            if not cached:
                self._kv_cache = None
            if ops.xattn_block_supported(x_ln, kv, heads=self.heads, n_latents=n):
                # to_q -> masked attention -> to_out in ONE cluster kernel (K1-fused)
                return ops.xattn_block(x_ln, self.to_q.weight, kv, text_time, self.to_out.weight,
                                       heads=self.heads, n_latents=n, scale=self.scale)
            q = ops.linear_acc(x_ln, self.to_q.weight)
            out = ops.masked_cross_attention(q, kv, text_time, heads=self.heads, n_latents=n,
                                             scale=self.scale)
            return ops.linear_acc(out, self.to_out.weight)
        if cached and T == 1:
            # keyed on the media tensor (identity + version) and on the projection weights'
            # version: new images of the same batch size, or an optimizer step between two
            # cached forwards, rebuild the cache (upstream recomputes to_kv(media) every call)
            kv = self.cached_media_kv(media)
            q = ops.linear_rows(x_ln, self.to_q.weight)
            out = ops.xattn_decode(q, kv, text_time[:, 0].contiguous(), heads=self.heads,
                                   n_latents=n, scale=self.scale)
            return ops.linear_rows(out, self.to_out.weight)
        q = ops.linear_acc(x_ln, self.to_q.weight)
        if cached:
            kv = self.cached_media_kv(media)
        else:
            self._kv_cache = None
            kv = self.project_media(media)
        out = ops.masked_cross_attention(q, kv, text_time, heads=self.heads, n_latents=n,
                                         scale=self.scale)                  # K1
        return ops.linear_acc(out, self.to_out.weight)


class GatedCrossAttentionBlock(nn.Module):
    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, ff_mult=4,
                 only_attend_immediate_media=True):
        super().__init__()
        self.attn = MaskedCrossAttention(dim=dim, dim_visual=dim_visual, dim_head=dim_head,
                                         heads=heads,
                                         only_attend_immediate_media=only_attend_immediate_media)
        self.attn_gate = nn.Parameter(torch.tensor([0.0]))
        self.ff = FeedForward(dim, mult=ff_mult)
        self.ff_gate = nn.Parameter(torch.tensor([0.0]))

    def forward(self, x, media, media_locations=None, use_cached_media=False, text_time=None,
                next_ln=None, x_ln=None):
        """`next_ln`: the LayerNorm module that consumes the block's output next (the decoder
        layer's input_layernorm); if given, its output is produced by the same K5 launch as the
        last gated residual and the call returns (x, next_ln(x))."""
        a = self.attn(x, media, media_locations=media_locations,
                      use_cached_media=use_cached_media, text_time=text_time, x_ln=x_ln)
        # x = a*tanh(attn_gate) + x, fused with ff's LayerNorm                     (K5)
        x, h = ops.gate_residual_ln(a, x, self.attn_gate, self.ff[0].weight, self.ff[0].bias,
                                    self.ff[0].eps)
        h = _ff_tail(self.ff, h)
        # x = ff*tanh(ff_gate) + x                                                 (K5)
        if next_ln is not None:
            return ops.gate_residual_ln(h, x, self.ff_gate, next_ln.weight, next_ln.bias, next_ln.eps)
        return ops.gate_residual(h, x, self.ff_gate)
