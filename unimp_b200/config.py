"""Shapes of the configurations BASELINE.json names.

Token-id layout for the 4B-instruct vocabulary follows reference `UniMP/mmrec.py:538-581`
(tokens are appended to the GPT-NeoX tokenizer in this order) and upstream's factory
(`<|endofchunk|>`, `<image>`, `<PAD>` first).  There is no tokenizer in this image, so ids
are assigned positionally.
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class SpecialTokens:
    bos: int
    eos: int
    endofchunk: int
    media: int
    pad: int
    answer: int
    first_item: int
    n_items: int
    first_img: int
    n_img: int
    n_plain: int  # ids [1, n_plain) are ordinary text


@dataclass
class FlamingoConfig:
    name: str
    # vision (CLIP ViT)
    vis_width: int
    vis_layers: int
    vis_heads: int
    vis_mlp: int
    image_size: int
    patch_size: int
    # language model (GPT-NeoX)
    lm_hidden: int
    lm_layers: int
    lm_heads: int
    lm_ffn: int
    vocab: int
    rotary_pct: float = 1.0
    use_parallel_residual: bool = False
    max_positions: int = 2048
    cross_attn_every_n_layers: int = 1
    # perceiver / x-attn keep upstream defaults (Flamingo.__init__ passes only `dim`)
    perceiver_depth: int = 6
    n_latents: int = 64
    xattn_heads: int = 8
    xattn_dim_head: int = 64
    ff_mult: int = 4
    tokens: SpecialTokens = field(default=None)

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2


def tiny_config() -> FlamingoConfig:
    """configs[0]: tiny random-init Flamingo, CPU-runnable."""
    V = 512
    tok = SpecialTokens(bos=0, eos=0, endofchunk=400, media=401, pad=402, answer=403,
                        first_item=404, n_items=64, first_img=468, n_img=44, n_plain=400)
    return FlamingoConfig(
        name="tiny", vis_width=64, vis_layers=2, vis_heads=4, vis_mlp=256, image_size=56,
        patch_size=14, lm_hidden=128, lm_layers=2, lm_heads=4, lm_ffn=512, vocab=V,
        cross_attn_every_n_layers=1, tokens=tok,
    )


def openflamingo_4b_config() -> FlamingoConfig:
    """configs[1..4]: OpenFlamingo-4B-instruct = ViT-L/14 + RedPajama-INCITE-Instruct-3B,
    x-attn every 2 layers (reference `UniMP/mmrec.py:505-514`), vocabulary grown as
    `UniMP/mmrec.py:538-595` does for subset "all": 50 277 + 3 + 1 + 5 + 5 + 22 738 + 1 024."""
    base = 50277
    tok = SpecialTokens(bos=0, eos=0, endofchunk=base, media=base + 1, pad=base + 2,
                        answer=base + 3, first_item=base + 14, n_items=22738,
                        first_img=base + 14 + 22738, n_img=1024, n_plain=base)
    V = base + 14 + 22738 + 1024
    assert V == 74053
    return FlamingoConfig(
        name="openflamingo-4b-instruct", vis_width=1024, vis_layers=24, vis_heads=16,
        vis_mlp=4096, image_size=224, patch_size=14, lm_hidden=2560, lm_layers=32,
        lm_heads=32, lm_ffn=10240, vocab=V, cross_attn_every_n_layers=2, tokens=tok,
    )


@dataclass
class Workload:
    name: str
    B: int
    Ti: int
    T: int
    gamma: float = 2.0
    row_weights: tuple = (2.0,)
    img_gen: bool = False


WORKLOADS = {
    # SURVEY §8d
    "C1-tiny": Workload("C1-tiny", B=2, Ti=2, T=32),
    "C2-rec": Workload("C2-rec", B=3, Ti=2, T=256),
    "C3-multitask": Workload("C3-multitask", B=3, Ti=8, T=1024, row_weights=(2.0, 1.0, 1.0, 1.0)),
    "C5-imggen": Workload("C5-imggen", B=3, Ti=2, T=1024, row_weights=(1.0,), img_gen=True),
}
