"""Plain-PyTorch restatement of the open_flamingo v2.0.1 model arithmetic (oracle).

TEST INFRASTRUCTURE — see oracle/__init__.py.  PARITY UNPINNED for the open_flamingo classes
below (the package is absent); the vision tower this file wraps (HF `CLIPVisionModel`) IS pinned
against the reference's in-tree `UniMP/xformers_model/clip.py` (tests/test_reference_golden.py).

The reference imports these classes from the un-vendored ``open_flamingo`` package
(reference ``UniMP/mmrec.py:20-22``; pinned ``requirements.txt:35``); the call sites
this file must serve are ``UniMP/mmrec.py:177-181`` (forward),
``UniMP/pipeline/eval/eval_rec.py:100-110`` (generate) and
``UniMP/mmrec.py:506-512`` (factory).  The arithmetic follows SURVEY.md §9.

Everything here is dense, eager and materialises the (B,h,T,Ti*n) similarity tensor
exactly as upstream does.  That is the point: it is the slow, obviously-faithful form
the CUDA kernels are compared against.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from einops import rearrange, repeat


def exists(v):
    return v is not None


def FeedForward(dim: int, mult: int = 4) -> nn.Sequential:
    """Sequential(LN, Linear(dim,4dim,no bias), GELU, Linear(4dim,dim,no bias))."""
    inner = int(dim * mult)
    return nn.Sequential(
        nn.LayerNorm(dim),
        nn.Linear(dim, inner, bias=False),
        nn.GELU(),
        nn.Linear(inner, dim, bias=False),
    )


class PerceiverAttention(nn.Module):
    """SURVEY §9 `PerceiverAttention`; state-dict keys norm_media/norm_latents/to_q/to_kv/to_out."""

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

    def forward(self, x, latents):
        # x (b,T,n1,D) image tokens; latents (b,T,n2,D)
        x = self.norm_media(x)
        latents = self.norm_latents(latents)
        h = self.heads
        q = self.to_q(latents)
        kv_input = torch.cat((x, latents), dim=-2)
        k, v = self.to_kv(kv_input).chunk(2, dim=-1)
        q, k, v = (rearrange(t, "b t n (h d) -> b h t n d", h=h) for t in (q, k, v))
        q = q * self.scale
        sim = torch.einsum("... i d, ... j d -> ... i j", q, k)
        sim = sim - sim.amax(dim=-1, keepdim=True).detach()
        attn = sim.softmax(dim=-1)
        out = torch.einsum("... i j, ... j d -> ... i d", attn, v)
        out = rearrange(out, "b h t n d -> b t n (h d)", h=h)
        return self.to_out(out)


class PerceiverResampler(nn.Module):
    """SURVEY §9 `PerceiverResampler` at the defaults Flamingo.__init__ uses (only `dim`)."""

    def __init__(self, *, dim, depth=6, dim_head=64, heads=8, num_latents=64, ff_mult=4):
        super().__init__()
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(
                nn.ModuleList(
                    [
                        PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads),
                        FeedForward(dim=dim, mult=ff_mult),
                    ]
                )
            )
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        # x (b, T, F, v, D) -> (b, T, n, D)
        b, T, F, v = x.shape[:4]
        x = rearrange(x, "b T F v d -> b T (F v) d")
        latents = repeat(self.latents, "n d -> b T n d", b=b, T=T)
        for attn, ff in self.layers:
            latents = attn(x, latents) + latents
            latents = ff(latents) + latents
        return self.norm(latents)


class MaskedCrossAttention(nn.Module):
    """SURVEY §9 `MaskedCrossAttention` (dense, materialised mask)."""

    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, only_attend_immediate_media=True):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner = dim_head * heads
        self.norm = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim_visual, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        self.only_attend_immediate_media = only_attend_immediate_media

    def forward(self, x, media, media_locations=None, use_cached_media=False):
        if not use_cached_media:
            assert media_locations.shape[1] == x.shape[1], (
                f"media_location.shape is {media_locations.shape} but x.shape is {x.shape}"
            )
        T_txt = x.shape[1]
        _, T_img, n = media.shape[:3]
        h = self.heads
        x = self.norm(x)
        q = self.to_q(x)
        media = rearrange(media, "b t n d -> b (t n) d")
        k, v = self.to_kv(media).chunk(2, dim=-1)
        q, k, v = (rearrange(t, "b n (h d) -> b h n d", h=h) for t in (q, k, v))
        q = q * self.scale
        sim = torch.einsum("... i d, ... j d -> ... i j", q, k)
        if exists(media_locations):
            media_time = torch.arange(T_img, device=x.device) + 1
            if use_cached_media:
                # text_time = number of media tokens in the cached prompt, for every new token
                text_time = repeat(
                    torch.count_nonzero(media_locations, dim=1), "b -> b i", i=T_txt
                )
            else:
                text_time = media_locations.cumsum(dim=-1)
            mask_op = torch.eq if self.only_attend_immediate_media else torch.ge
            text_to_media_mask = mask_op(
                rearrange(text_time, "b i -> b 1 i 1"),
                repeat(media_time, "j -> 1 1 1 (j n)", n=n),
            )
            sim = sim.masked_fill(~text_to_media_mask, -torch.finfo(sim.dtype).max)
        sim = sim - sim.amax(dim=-1, keepdim=True).detach()
        attn = sim.softmax(dim=-1)
        if exists(media_locations) and self.only_attend_immediate_media:
            text_without_media_mask = text_time == 0
            text_without_media_mask = rearrange(text_without_media_mask, "b i -> b 1 i 1")
            attn = attn.masked_fill(text_without_media_mask, 0.0)
        out = torch.einsum("... i j, ... j d -> ... i d", attn, v)
        out = rearrange(out, "b h n d -> b n (h d)")
        return self.to_out(out)


class GatedCrossAttentionBlock(nn.Module):
    """SURVEY §9 `GatedCrossAttentionBlock`: tanh-gated x-attn then tanh-gated FF."""

    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, ff_mult=4,
                 only_attend_immediate_media=True):
        super().__init__()
        self.attn = MaskedCrossAttention(
            dim=dim, dim_visual=dim_visual, dim_head=dim_head, heads=heads,
            only_attend_immediate_media=only_attend_immediate_media,
        )
        self.attn_gate = nn.Parameter(torch.tensor([0.0]))
        self.ff = FeedForward(dim, mult=ff_mult)
        self.ff_gate = nn.Parameter(torch.tensor([0.0]))

    def forward(self, x, media, media_locations=None, use_cached_media=False):
        x = self.attn(x, media, media_locations=media_locations,
                      use_cached_media=use_cached_media) * self.attn_gate.tanh() + x
        x = self.ff(x) * self.ff_gate.tanh() + x
        return x


class FlamingoLayer(nn.Module):
    """SURVEY §9 `FlamingoLayer`: optional gated x-attn, then the (frozen) decoder layer."""

    def __init__(self, gated_cross_attn_layer, decoder_layer):
        super().__init__()
        self.gated_cross_attn_layer = gated_cross_attn_layer
        self.decoder_layer = decoder_layer
        self.vis_x = None
        self.media_locations = None
        self.use_cached_media = False

    def is_conditioned(self) -> bool:
        return self.vis_x is not None and self.media_locations is not None

    def condition_vis_x(self, vis_x):
        self.vis_x = vis_x

    def condition_media_locations(self, media_locations):
        self.media_locations = media_locations

    def condition_use_cached_media(self, use_cached_media):
        self.use_cached_media = use_cached_media

    def forward(self, lang_x, attention_mask=None, **decoder_layer_kwargs):
        if self.gated_cross_attn_layer is not None:
            if self.vis_x is None:
                raise ValueError("vis_x must be conditioned before forward pass")
            if self.media_locations is None:
                raise ValueError("media_locations must be conditioned before forward pass")
            lang_x = self.gated_cross_attn_layer(
                lang_x, self.vis_x, media_locations=self.media_locations,
                use_cached_media=self.use_cached_media,
            )
        return self.decoder_layer(lang_x, attention_mask=attention_mask, **decoder_layer_kwargs)


class FlamingoLM(nn.Module):
    """Oracle stand-in for `FlamingoLMMixin` grafted on HF GPTNeoXForCausalLM.

    Upstream mutates the HF instance's class; the oracle wraps it instead, which is
    arithmetically identical: decoder layers are replaced by FlamingoLayer, forward
    derives ``media_locations = input_ids == media_token_id`` and conditions every layer.
    """

    def __init__(self, lang_encoder, *, media_token_id, lang_hidden_size, vis_hidden_size,
                 cross_attn_every_n_layers):
        super().__init__()
        self.lm = lang_encoder
        self.media_token_id = media_token_id
        layers = self.lm.gpt_neox.layers
        self.gated_cross_attn_layers = nn.ModuleList(
            [
                GatedCrossAttentionBlock(dim=lang_hidden_size, dim_visual=vis_hidden_size)
                if (i + 1) % cross_attn_every_n_layers == 0 else None
                for i in range(len(layers))
            ]
        )
        self.lm.gpt_neox.layers = nn.ModuleList(
            [FlamingoLayer(g, d) for g, d in zip(self.gated_cross_attn_layers, layers)]
        )
        self._use_cached_vision_x = False

    @property
    def layers(self):
        return self.lm.gpt_neox.layers

    def is_conditioned(self):
        return all(l.is_conditioned() for l in self.layers)

    def clear_conditioned_layers(self):
        for l in self.layers:
            l.condition_vis_x(None)
            l.condition_media_locations(None)
            l.condition_use_cached_media(None)

    def forward(self, input_ids, attention_mask=None, **kw):
        media_locations = input_ids == self.media_token_id
        use_cached = (
            self._use_cached_vision_x and self.is_conditioned() and not media_locations.any()
        )
        for l in self.layers:
            if not use_cached:
                l.condition_media_locations(media_locations)
            l.condition_use_cached_media(use_cached)
        return self.lm(input_ids=input_ids, attention_mask=attention_mask, **kw)


class Flamingo(nn.Module):
    """SURVEY §9 `Flamingo`: vision encoder (frozen, no_grad) -> perceiver -> conditioned LM."""

    def __init__(self, vision_encoder, lang_encoder, eoc_token_id, media_token_id, vis_dim,
                 cross_attn_every_n_layers=1):
        super().__init__()
        self.eoc_token_id = eoc_token_id
        self.media_token_id = media_token_id
        self.vis_dim = vis_dim
        self.vision_encoder = vision_encoder  # HF CLIPVisionModel
        self.perceiver = PerceiverResampler(dim=vis_dim)
        self.lang_encoder = FlamingoLM(
            lang_encoder, media_token_id=media_token_id,
            lang_hidden_size=lang_encoder.config.hidden_size, vis_hidden_size=vis_dim,
            cross_attn_every_n_layers=cross_attn_every_n_layers,
        )
        self._use_cached_vision_x = False

    def forward(self, vision_x, lang_x, attention_mask=None, labels=None,
                clear_conditioned_layers=True, past_key_values=None, use_cache=False):
        assert (
            self.lang_encoder.is_conditioned() or vision_x is not None
        ) or self._use_cached_vision_x, "Must provide vision_x or have precached media"
        if self._use_cached_vision_x:
            assert vision_x is None
            assert self.lang_encoder.is_conditioned()
        else:
            self._encode_vision_x(vision_x)
        out = self.lang_encoder(
            input_ids=lang_x, attention_mask=attention_mask, labels=labels,
            past_key_values=past_key_values, use_cache=use_cache,
        )
        if clear_conditioned_layers:
            self.lang_encoder.clear_conditioned_layers()
        return out

    def _encode_vision_x(self, vision_x):
        assert vision_x.ndim == 6, "vision_x should be of shape (b, T_img, F, C, H, W)"
        b, T, F = vision_x.shape[:3]
        assert F == 1, "Only single frame supported"
        vision_x = rearrange(vision_x, "b T F c h w -> (b T F) c h w")
        with torch.no_grad():
            # open_clip `visual(x)[1]` with output_tokens=True == patch tokens, no CLS,
            # before ln_post == HF last_hidden_state[:, 1:]
            tokens = self.vision_encoder(pixel_values=vision_x).last_hidden_state[:, 1:]
        vision_x = rearrange(tokens, "(b T F) v d -> b T F v d", b=b, T=T, F=F)
        vision_x = self.perceiver(vision_x)
        for layer in self.lang_encoder.layers:
            layer.condition_vis_x(vision_x)
        return vision_x

    def cache_media(self, input_ids, vision_x):
        """Upstream `cache_media`: encode once and remember media_locations for decode."""
        self._encode_vision_x(vision_x)
        media_locations = input_ids == self.media_token_id
        for l in self.lang_encoder.layers:
            l.condition_media_locations(media_locations)
        self._use_cached_vision_x = True
        self.lang_encoder._use_cached_vision_x = True

    def uncache_media(self):
        self.lang_encoder.clear_conditioned_layers()
        self._use_cached_vision_x = False
        self.lang_encoder._use_cached_vision_x = False


def build_oracle_flamingo(vision_cfg, lm_cfg, *, media_token_id, eoc_token_id,
                          cross_attn_every_n_layers=1, seed=0, gate=0.5):
    """Random-init oracle model (there is no network for weights; BASELINE.json says so).

    Mirrors upstream `create_model_and_transforms` freezing: everything frozen except
    perceiver, gated_cross_attn_layers and the LM *input* embeddings.  `gate` sets
    attn_gate/ff_gate (upstream init 0 makes the x-attn branch vanish).
    """
    from transformers import CLIPVisionModel, GPTNeoXForCausalLM

    torch.manual_seed(seed)
    vis = CLIPVisionModel(vision_cfg)
    lm = GPTNeoXForCausalLM(lm_cfg)
    model = Flamingo(vis, lm, eoc_token_id, media_token_id, vis_dim=vision_cfg.hidden_size,
                     cross_attn_every_n_layers=cross_attn_every_n_layers)
    model.requires_grad_(False)
    model.perceiver.requires_grad_(True)
    model.lang_encoder.gated_cross_attn_layers.requires_grad_(True)
    model.lang_encoder.lm.get_input_embeddings().requires_grad_(True)
    with torch.no_grad():
        for blk in model.lang_encoder.gated_cross_attn_layers:
            if blk is not None:
                blk.attn_gate.fill_(gate)
                blk.ff_gate.fill_(gate)
    return model
