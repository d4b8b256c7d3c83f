"""Shared helpers for the parity tests: build the oracle and the product with equal weights."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def hf_configs(cfg):
    from transformers import CLIPVisionConfig, GPTNeoXConfig

    vc = CLIPVisionConfig(hidden_size=cfg.vis_width, num_hidden_layers=cfg.vis_layers,
                          num_attention_heads=cfg.vis_heads, intermediate_size=cfg.vis_mlp,
                          image_size=cfg.image_size, patch_size=cfg.patch_size,
                          hidden_act="quick_gelu")
    lc = GPTNeoXConfig(hidden_size=cfg.lm_hidden, num_hidden_layers=cfg.lm_layers,
                       num_attention_heads=cfg.lm_heads, intermediate_size=cfg.lm_ffn,
                       vocab_size=cfg.vocab, rotary_pct=cfg.rotary_pct,
                       use_parallel_residual=cfg.use_parallel_residual,
                       max_position_embeddings=cfg.max_positions, tie_word_embeddings=False,
                       hidden_dropout=0.0, attention_dropout=0.0)
    return vc, lc


def build_oracle(cfg, seed=0, gate=0.5):
    from oracle.flamingo_oracle import build_oracle_flamingo

    vc, lc = hf_configs(cfg)
    m = build_oracle_flamingo(vc, lc, media_token_id=cfg.tokens.media,
                              eoc_token_id=cfg.tokens.endofchunk,
                              cross_attn_every_n_layers=cfg.cross_attn_every_n_layers, seed=seed,
                              gate=gate)
    return m.float()


def copy_oracle_weights(oracle, product):
    """oracle (HF CLIP + wrapped HF GPT-NeoX) -> product (own ViT + mixin-grafted HF GPT-NeoX)."""
    from unimp_b200.vit import load_hf_clip_vision_weights

    dt = next(product.parameters()).dtype
    dev = next(product.parameters()).device

    def cast(sd):
        return {k: v.to(device=dev, dtype=dt if v.is_floating_point() else v.dtype)
                for k, v in sd.items()}

    load_hf_clip_vision_weights(product.vision_encoder, cast(oracle.vision_encoder.state_dict()))
    product.perceiver.load_state_dict(cast(oracle.perceiver.state_dict()), strict=True)
    missing, unexpected = product.lang_encoder.load_state_dict(
        cast(oracle.lang_encoder.lm.state_dict()), strict=False)
    assert not unexpected, unexpected
    # only the duplicate registrations may be "missing" (they alias tensors that were loaded)
    assert all(k.startswith(("gated_cross_attn_layers.", "old_decoder_blocks.")) for k in missing), missing
    return product


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def worst_elem(a, b):
    """Largest single-element error, each element measured against |reference element| + the
    reference tensor's RMS (relative for large elements, RMS-absolute for small ones: gradients
    are heavy-tailed).  A Frobenius-relative bound lets one wrong row of a 1536-row tensor
    through; this does not (a wrong row scores ~0.5-1)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rms = b.pow(2).mean().sqrt().clamp_min(1e-30)
    return float(((a - b).abs() / (b.abs() + rms)).max())


def assert_close(a, b, tol, what="", elem_factor=5.0):
    """Frobenius-relative error < tol AND every element within elem_factor * tol of
    (|reference element| + reference RMS)  (bf16: tol 2e-2 -> 0.1; fp32: tol 1e-4 -> 5e-4)."""
    fro, worst = rel_err(a, b), worst_elem(a, b)
    assert fro < tol, f"{what}: Frobenius-relative error {fro:.3e} >= {tol:.1e}"
    assert worst < elem_factor * tol, f"{what}: worst element off by {worst:.3e} of (|ref| + RMS) >= {elem_factor * tol:.1e}"


def max_rel(a, b, floor=1e-6):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())
