"""create_model_and_transforms — the factory the reference calls (`UniMP/mmrec.py:506-512`).

Upstream downloads CLIP ViT-L/14 and a HF causal LM; there is no network here
(BASELINE.json: random-init weights of the named architecture), so `lang_encoder_path` /
`clip_vision_encoder_path` select an ARCHITECTURE and weights are random-initialised; real
OpenFlamingo / UniMP `.pt` checkpoints (trainable params only, reference
`UniMP/pipeline/train/train_utils.py:258-265`) load afterwards with
`model.load_state_dict(sd, strict=False)` exactly as `UniMP/mmrec.py:513-514` does, because
state-dict key names are upstream's.
"""
from __future__ import annotations

import torch
from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

from .config import FlamingoConfig, openflamingo_4b_config, tiny_config
from .flamingo import Flamingo
from .flamingo_lm import FlamingoLMMixin, extend_instance, wrap_output_head
from .vit import VisionTransformer

_LM_ARCH = {
    "togethercomputer/RedPajama-INCITE-Instruct-3B-v1": openflamingo_4b_config,
    "togethercomputer/RedPajama-INCITE-Base-3B-v1": openflamingo_4b_config,
    "tiny": tiny_config,
}


def _lm_config(cfg: FlamingoConfig) -> GPTNeoXConfig:
    c = GPTNeoXConfig(
        hidden_size=cfg.lm_hidden, num_hidden_layers=cfg.lm_layers,
        num_attention_heads=cfg.lm_heads, intermediate_size=cfg.lm_ffn, vocab_size=cfg.vocab,
        rotary_pct=cfg.rotary_pct, use_parallel_residual=cfg.use_parallel_residual,
        max_position_embeddings=cfg.max_positions, tie_word_embeddings=False,
        hidden_dropout=0.0, attention_dropout=0.0)
    c._attn_implementation = "sdpa"
    return c


def build_flamingo(cfg: FlamingoConfig, *, dtype=torch.bfloat16, device="cuda", seed=0,
                   gate=None):
    """Random-init Flamingo of architecture `cfg` on `device` in `dtype`, frozen like upstream
    (everything frozen except perceiver, gated_cross_attn_layers, LM input embeddings)."""
    torch.manual_seed(seed)
    with torch.device(device):
        vis = VisionTransformer(cfg.image_size, cfg.patch_size, cfg.vis_width, cfg.vis_layers,
                                cfg.vis_heads, cfg.vis_mlp)
        lm = GPTNeoXForCausalLM(_lm_config(cfg))
    extend_instance(lm, FlamingoLMMixin)
    lm.set_decoder_layers_attr_name("gpt_neox.layers")
    with torch.device(device):
        model = Flamingo(vis, lm, cfg.tokens.endofchunk, cfg.tokens.media, vis_dim=cfg.vis_width,
                         cross_attn_every_n_layers=cfg.cross_attn_every_n_layers)
    model.to(dtype)
    model.requires_grad_(False)
    model.perceiver.requires_grad_(True)
    model.lang_encoder.gated_cross_attn_layers.requires_grad_(True)
    model.lang_encoder.get_input_embeddings().requires_grad_(True)
    wrap_output_head(model.lang_encoder)
    if gate is not None:
        with torch.no_grad():
            for blk in model.lang_encoder.gated_cross_attn_layers:
                if blk is not None:
                    blk.attn_gate.fill_(gate)
                    blk.ff_gate.fill_(gate)
    return model


class SyntheticTokenizer:
    """Tokenizer stand-in carrying the special-token ids the train loop looks up
    (reference `UniMP/mmrec.py:84-88`); there are no tokenizer files in this image."""

    def __init__(self, cfg: FlamingoConfig):
        t = cfg.tokens
        self.vocab = {"<image>": t.media, "<|endofchunk|>": t.endofchunk, "<answer>": t.answer,
                      "<PAD>": t.pad}
        self.pad_token_id, self.eos_token_id, self.bos_token_id = t.pad, t.eos, t.bos
        self._len = cfg.vocab

    def __call__(self, text, add_special_tokens=False):
        return {"input_ids": [self.vocab[text]]}

    def __len__(self):
        return self._len


def create_model_and_transforms(clip_vision_encoder_path: str, clip_vision_encoder_pretrained: str,
                                lang_encoder_path: str, tokenizer_path: str,
                                cross_attn_every_n_layers: int = 1, *, dtype=torch.bfloat16,
                                device="cuda", seed=0, **_ignored):
    """Same signature and return triple as upstream: (model, image_processor, tokenizer)."""
    if lang_encoder_path not in _LM_ARCH:
        raise ValueError(f"unknown lang_encoder_path {lang_encoder_path!r}; known: {list(_LM_ARCH)}")
    cfg = _LM_ARCH[lang_encoder_path]()
    cfg.cross_attn_every_n_layers = cross_attn_every_n_layers
    model = build_flamingo(cfg, dtype=dtype, device=device, seed=seed)
    image_processor = None  # CLIP preprocessing is CPU/PIL-side data prep (out)
    return model, image_processor, SyntheticTokenizer(cfg)
