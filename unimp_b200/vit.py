"""CLIP ViT vision tower (frozen, forward only) on the sm_100a attention / LayerNorm kernels.

Upstream uses open_clip's `VisionTransformer` with `output_tokens=True`:
Flamingo consumes `visual(x)[1]` = the patch tokens (no CLS) BEFORE `ln_post`.  Parameter names
follow open_clip (`conv1`, `class_embedding`, `positional_embedding`, `ln_pre`,
`transformer.resblocks.{i}.{ln_1,attn.in_proj_*,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj}`,
`ln_post`) so an open_clip ViT-L/14 state dict loads directly.  Self-attention (K3, 257x257,
16 heads x 64) is `unimp_attn_fwd`; residual+LayerNorm pairs are `unimp_gate_residual_ln_fwd`.
The dead `UniMP/xformers_model/clip.py:130-136` attention call has the same shape.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _Attn(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        self.heads = heads
        self.width = width


class _Mlp(nn.Module):
    def __init__(self, width, mlp):
        super().__init__()
        self.c_fc = nn.Linear(width, mlp)
        self.c_proj = nn.Linear(mlp, width)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, width, heads, mlp):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = _Attn(width, heads)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = _Mlp(width, mlp)


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, mlp)
                                        for _ in range(layers)])


class VisionTransformer(nn.Module):
    def __init__(self, image_size=224, patch_size=14, width=1024, layers=24, heads=16, mlp=4096,
                 quick_gelu=True):
        super().__init__()
        assert width % heads == 0 and width // heads <= 64, "head dim must be <= 64"
        # head dim 64 (ViT-L/14) runs on the attention kernels directly; a smaller head dim (the
        # tiny CPU-runnable configuration of BASELINE.md: 64-wide, 4 heads x 16) is zero-padded
        # to 64 per head — the extra coordinates add 0 to every score and their outputs are dropped
        self.image_size, self.patch_size, self.width, self.heads = image_size, patch_size, width, heads
        self.grid = image_size // patch_size
        self.conv1 = nn.Conv2d(3, width, patch_size, patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads, mlp)
        self.ln_post = nn.LayerNorm(width)
        self.quick_gelu = quick_gelu
        self.output_tokens = True
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=scale)

    def _attention(self, qkv):
        W, H = self.width, self.heads
        dh = W // H
        if dh == 64:
            return ops.attention(qkv[..., :W], qkv[..., W:], heads=H, scale=0.125)
        N, L, _ = qkv.shape
        q, k, v = (F.pad(t, (0, 64 - dh)).reshape(N, L, H * 64)
                   for t in qkv.view(N, L, 3, H, dh).unbind(2))
        a = ops.attention(q, torch.cat((k, v), dim=-1), heads=H, scale=dh ** -0.5)
        return a.view(N, L, H, 64)[..., :dh].reshape(N, L, W)

    def _act(self, h):
        return ops.quick_gelu_(h) if self.quick_gelu else F.gelu(h)

    @torch.no_grad()
    def forward(self, x):
        """x (N,3,H,W) -> (pooled placeholder None, tokens (N, grid^2, width)) like
        open_clip's `visual(x)` with output_tokens=True; index [1] is what Flamingo uses."""
        N = x.shape[0]
        x = self.conv1(x.to(self.conv1.weight.dtype))
        x = x.reshape(N, self.width, -1).permute(0, 2, 1)
        cls = self.class_embedding.to(x.dtype).expand(N, 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = ops.layer_norm(x.contiguous(), self.ln_pre.weight, self.ln_pre.bias, self.ln_pre.eps)
        blocks = self.transformer.resblocks
        b0 = blocks[0]
        h = ops.layer_norm(x, b0.ln_1.weight, b0.ln_1.bias, b0.ln_1.eps)
        W = self.width
        for i, blk in enumerate(blocks):
            qkv = F.linear(h, blk.attn.in_proj_weight, blk.attn.in_proj_bias)  # (N, L, 3W)
            a = self._attention(qkv)                                                  # K3
            a = blk.attn.out_proj(a)
            x, h = ops.gate_residual_ln(a, x, None, blk.ln_2.weight, blk.ln_2.bias, blk.ln_2.eps)
            m = blk.mlp.c_proj(self._act(blk.mlp.c_fc(h)))
            if i + 1 < len(blocks):
                nb = blocks[i + 1]
                x, h = ops.gate_residual_ln(m, x, None, nb.ln_1.weight, nb.ln_1.bias, nb.ln_1.eps)
            else:
                x = ops.gate_residual(m, x, None)
        return None, x[:, 1:]


def load_hf_clip_vision_weights(vit: VisionTransformer, hf_state: dict):
    """Copy a HF `CLIPVisionModel` state dict (the oracle's vision tower) into `vit`."""
    p = "vision_model."
    sd = {}
    sd["class_embedding"] = hf_state[p + "embeddings.class_embedding"]
    sd["conv1.weight"] = hf_state[p + "embeddings.patch_embedding.weight"]
    sd["positional_embedding"] = hf_state[p + "embeddings.position_embedding.weight"]
    for a, b in (("ln_pre", "pre_layrnorm"), ("ln_post", "post_layernorm")):
        sd[f"{a}.weight"] = hf_state[p + f"{b}.weight"]
        sd[f"{a}.bias"] = hf_state[p + f"{b}.bias"]
    for i in range(len(vit.transformer.resblocks)):
        s = p + f"encoder.layers.{i}."
        d = f"transformer.resblocks.{i}."
        sd[d + "attn.in_proj_weight"] = torch.cat(
            [hf_state[s + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        sd[d + "attn.in_proj_bias"] = torch.cat(
            [hf_state[s + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"),
                     ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            sd[d + a + ".weight"] = hf_state[s + b + ".weight"]
            sd[d + a + ".bias"] = hf_state[s + b + ".bias"]
    missing, unexpected = vit.load_state_dict(sd, strict=True)
    return vit
