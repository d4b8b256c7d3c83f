"""Synthetic batches with the shapes and prompt structure of the reference's rec data path.

There is no dataset or tokenizer in this image (BASELINE.json: synthetic data, random-init
weights), so this reproduces only the *layout* the reference's collate produces:
`RecDataset.process_train_rec_pair` builds, per history item,
``"<image> {meta} <answer> item_{id} <|endofchunk|> "`` followed by a question and
``<answer> item_{next}`` (reference `UniMP/pipeline/mm_utils/rec_dataset.py:414,424`), wraps
BOS/EOS (`:444-445`), and `collate_fn` right-pads ids with pad_token_id / masks with 0 and
stacks images (`UniMP/pipeline/mm_utils/collate_rec.py:51-55,70-72`).  Output keys follow
`collate_rec.py:59-72`: input_ids, attention_masks, patch_images, weights.
`tests/test_reference_golden.py` checks the rows against the token grammar of batches built by the
reference's own `RecDataset.process_train_rec_pair` + `collate_fn` (fixture
`tests/golden/ref_dataset_batches.pt`).
"""
from __future__ import annotations

import torch

from .config import FlamingoConfig, Workload


def make_batch(cfg: FlamingoConfig, wl: Workload, *, seed: int = 1234, ragged: bool = False,
               device="cpu", image_dtype=torch.float32):
    """Returns dict(input_ids (B,T) i64, attention_masks (B,T) i64, patch_images
    (B,Ti,3,H,W), weights (B,) f32).  `ragged=True` makes sample 0 carry fewer `<image>`
    tokens than Ti and start with text before its first `<image>` (SURVEY §8d variant)."""
    g = torch.Generator().manual_seed(seed)
    tk = cfg.tokens
    B, Ti, T = wl.B, wl.Ti, wl.T
    ids = torch.full((B, T), tk.pad, dtype=torch.int64)
    mask = torch.zeros((B, T), dtype=torch.int64)

    def rnd_text(n):
        return torch.randint(1, tk.n_plain, (n,), generator=g).tolist()

    def rnd_item():
        return int(torch.randint(tk.first_item, tk.first_item + tk.n_items, (1,), generator=g))

    def rnd_img_tokens(n):
        return torch.randint(tk.first_img, tk.first_img + tk.n_img, (n,), generator=g).tolist()

    for b in range(B):
        n_img = Ti
        lead = 0
        if ragged and b == 0:
            n_img = max(1, Ti - 1)
            lead = 3
        # budget: BOS + lead + n_img chunks + question + <answer> ans + EOS  <= T, leave pad
        fill = int(torch.randint(int(0.75 * T), T + 1, (1,), generator=g))
        ans_len = 257 if wl.img_gen else 1
        ans_len = min(ans_len, max(1, fill // 3))
        tail = 2 + ans_len  # <answer> ans.. EOS
        q_len = max(1, min(12, fill // 8))
        per_chunk = max(4, (fill - 1 - lead - q_len - tail) // max(n_img, 1))
        seq = [tk.bos] + rnd_text(lead)
        for _ in range(n_img):
            meta = max(1, per_chunk - 4)
            jitter = int(torch.randint(0, max(1, meta // 4) + 1, (1,), generator=g))
            seq += [tk.media] + rnd_text(max(1, meta - jitter)) + [tk.answer, rnd_item(), tk.endofchunk]
        seq += rnd_text(q_len) + [tk.answer]
        seq += rnd_img_tokens(ans_len) if wl.img_gen else [rnd_item()]
        seq += [tk.eos]
        seq = seq[:T]
        ids[b, : len(seq)] = torch.tensor(seq, dtype=torch.int64)
        mask[b, : len(seq)] = 1
    images = torch.randn((B, Ti, 3, cfg.image_size, cfg.image_size), generator=g).to(image_dtype)
    rw = wl.row_weights
    weights = torch.tensor([rw[b % len(rw)] for b in range(B)], dtype=torch.float32)
    return {
        "input_ids": ids.to(device),
        "attention_masks": mask.to(device),
        "patch_images": images.to(device),
        "weights": weights.to(device),
    }
