"""Flamingo — the module the reference trains (`open_flamingo.Flamingo`).

Call surface kept (reference `UniMP/mmrec.py:177-182,190`, `eval_rec.py:100-110`,
`mmrec_prefix.py:631-633`): `model(vision_x=, lang_x=, attention_mask=, labels=)` returning an
object indexable as `out[0]` / `out["logits"]`; `model.generate(vision_x=, lang_x=, ...)`;
attributes `.vision_encoder`, `.perceiver`, `.lang_encoder` (+ `.gated_cross_attn_layers`,
`.get_input_embeddings()`, `.resize_token_embeddings()`), `cache_media/uncache_media`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .helpers import PerceiverResampler


class Flamingo(nn.Module):
    def __init__(self, vision_encoder, lang_encoder, eoc_token_id, media_token_id, vis_dim,
                 cross_attn_every_n_layers=1, gradient_checkpointing=False):
        super().__init__()
        self.eoc_token_id = eoc_token_id
        self.media_token_id = media_token_id
        self.vis_dim = vis_dim
        if hasattr(lang_encoder.config, "d_model"):
            self.lang_dim = lang_encoder.config.d_model
        else:
            self.lang_dim = lang_encoder.config.hidden_size
        self.vision_encoder = vision_encoder
        self.perceiver = PerceiverResampler(dim=self.vis_dim)
        self.lang_encoder = lang_encoder
        self.lang_encoder.init_flamingo(
            media_token_id=media_token_id, lang_hidden_size=self.lang_dim,
            vis_hidden_size=self.vis_dim, cross_attn_every_n_layers=cross_attn_every_n_layers,
            gradient_checkpointing=gradient_checkpointing)
        self._use_cached_vision_x = False

    def forward(self, vision_x, lang_x, attention_mask=None, labels=None,
                clear_conditioned_layers=True, past_key_values=None, use_cache=False,
                label_rows=None):
        """`label_rows` is product-only (default: upstream behaviour): see FlamingoLMMixin.forward."""
        assert (
            self.lang_encoder.initialized_flamingo
        ), "Flamingo layers are not initialized. Please call `init_flamingo` first."
        assert (
            self.lang_encoder._use_cached_vision_x or vision_x is not None
        ), "Must provide either vision_x or have precached media using cache_media()."
        if self.lang_encoder._use_cached_vision_x:
            assert vision_x is None, (
                "Expect vision_x to be None when media has been cached using cache_media(). "
                "Try uncache_media() first.")
            assert self.lang_encoder.is_conditioned()
        else:
            self._encode_vision_x(vision_x=vision_x)
            self._condition_media_locations(input_ids=lang_x)
        extra = {} if label_rows is None else {"label_rows": label_rows}
        output = self.lang_encoder(input_ids=lang_x, attention_mask=attention_mask, labels=labels,
                                   past_key_values=past_key_values, use_cache=use_cache, **extra)
        if clear_conditioned_layers:
            self.lang_encoder.clear_conditioned_layers()
        return output

    @torch.no_grad()
    def generate(self, vision_x, lang_x, attention_mask=None, **kwargs):
        num_beams = kwargs.pop("num_beams", 1)
        if num_beams > 1:
            vision_x = vision_x.repeat_interleave(num_beams, dim=0)
        self.lang_encoder._use_cached_vision_x = True
        self._encode_vision_x(vision_x=vision_x)
        eos_token_id = kwargs.pop("eos_token_id", self.eoc_token_id)
        output = self.lang_encoder.generate(input_ids=lang_x, attention_mask=attention_mask,
                                            eos_token_id=eos_token_id, num_beams=num_beams,
                                            **kwargs)
        self.lang_encoder.clear_conditioned_layers()
        self.lang_encoder._use_cached_vision_x = False
        return output

    def _encode_vision_x(self, vision_x: torch.Tensor):
        assert vision_x.ndim == 6, "vision_x should be of shape (b, T_img, F, C, H, W)"
        b, T, F = vision_x.shape[:3]
        assert F == 1, "Only single frame supported"
        vision_x = vision_x.reshape(b * T * F, *vision_x.shape[3:])
        with torch.no_grad():
            tokens = self.vision_encoder(vision_x)[1]           # (bTF, v, d) patch tokens
        tokens = tokens.reshape(b, T, F, tokens.shape[-2], tokens.shape[-1])
        vision_x = self.perceiver(tokens)                        # (b, T, n, d)
        for layer in self.lang_encoder._get_decoder_layers():
            layer.condition_vis_x(vision_x)

    def _condition_media_locations(self, input_ids: torch.Tensor):
        media_locations = input_ids == self.media_token_id
        for layer in self.lang_encoder._get_decoder_layers():
            layer.condition_media_locations(media_locations)

    def cache_media(self, input_ids: torch.Tensor, vision_x: torch.Tensor):
        self._encode_vision_x(vision_x=vision_x)
        self._condition_media_locations(input_ids=input_ids)
        self.lang_encoder._use_cached_vision_x = True

    def uncache_media(self):
        self.lang_encoder.clear_conditioned_layers()
        self.lang_encoder._use_cached_vision_x = False
