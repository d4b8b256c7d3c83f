"""Generates tests/golden/*.pt from the oracle (the reference ships no fixtures: SURVEY §8c).

Run from the repo root:  python tests/golden/make_golden.py
The vectors pin (a) the oracle against drift and (b) the CUDA path on the GPU box, where
neither /root/reference nor a second implementation exists.  Small on purpose (< 300 KB).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from util import build_oracle  # noqa: E402

from oracle.loss_oracle import focal_loss, mask_labels  # noqa: E402
from unimp_b200 import tiny_config  # noqa: E402
from unimp_b200.config import WORKLOADS  # noqa: E402
from unimp_b200.synth import make_batch  # noqa: E402


def main():
    torch.set_num_threads(1)  # bit-stable reductions
    cfg = tiny_config()
    wseed, dseed, gate = 0, 1234, 0.5
    model = build_oracle(cfg, seed=wseed, gate=gate)
    batch = make_batch(cfg, WORKLOADS["C1-tiny"], seed=dseed, ragged=True)
    labels = mask_labels(batch["input_ids"], answer_token_id=cfg.tokens.answer,
                         endofchunk_token_id=cfg.tokens.endofchunk,
                         media_token_id=cfg.tokens.media, pad_token_id=cfg.tokens.pad)
    out = model(vision_x=batch["patch_images"].unsqueeze(2), lang_x=batch["input_ids"],
                attention_mask=batch["attention_masks"], labels=labels)
    loss = focal_loss(out.logits, labels, batch["weights"], gamma=2.0)
    loss.backward()
    blk = model.lang_encoder.gated_cross_attn_layers[0]
    grads = {
        "attn_gate0": blk.attn_gate.grad.clone(),
        "ff_gate0": blk.ff_gate.grad.clone(),
        "to_q0": blk.attn.to_q.weight.grad.clone(),
        "to_kv0": blk.attn.to_kv.weight.grad[:, :8].clone(),
        "latents": model.perceiver.latents.grad.clone(),
        "embed_in_rows": model.lang_encoder.lm.get_input_embeddings().weight.grad[
            batch["input_ids"][0, :8]].clone(),
    }
    torch.save({
        "weight_seed": wseed, "data_seed": dseed, "gate": gate,
        "input_ids": batch["input_ids"], "labels": labels,
        "logits": out.logits.detach(), "loss": loss.detach(), "hf_loss": out.loss.detach(),
        "grads": grads,
    }, os.path.join(HERE, "tiny_fwd_loss.pt"))
    print("loss", float(loss), "hf_loss", float(out.loss), "n_valid", int((labels[:, 1:] != -100).sum()))


if __name__ == "__main__":
    main()
