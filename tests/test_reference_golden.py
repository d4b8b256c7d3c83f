"""Parity against vectors produced by RUNNING THE REFERENCE'S OWN CODE.

`tests/golden/ref_*.pt` come from `tests/golden/make_reference_golden.py`, which imports the
unmodified reference from /root/reference (in the build container only) and drives
`UniMP/mmrec.py::train_one_epoch`, `collate_rec.py::collate_fn`, `train_utils.py::get_checkpoint`
and the ast-lifted `get_grouped_params`.  These pin the IN-TREE half of the path (labels, focal
loss and its gradient, batch layout, checkpoint / weight-decay rules):

* CPU (`-m "not gpu"`): the oracle and the host logic against the reference's outputs;
* GPU (`-m gpu`): the CUDA kernels, through the C ABI, against the same outputs.

Bars: labels / collate / key sets bit-exact; fp32 loss 1e-6 (oracle) and 1e-5 (CUDA), gradient
1e-6 / 1e-4 Frobenius-relative; bf16 loss 1e-3, gradient 2e-2 (north-star tolerances).
"""
import os

import pytest
import torch

from util import GOLDEN, rel_err

DEV = "cuda"


def _cases():
    blob = torch.load(os.path.join(GOLDEN, "ref_train_step.pt"), weights_only=False)
    return blob["tokens"], blob["cases"]


def _case_ids():
    return [c["name"] for c in _cases()[1]]


def _dense_grad(case):
    B, T = case["input_ids"].shape
    V = case["logits_i8"].shape[-1]
    d = torch.zeros(B * T, V)
    d[case["dlogits_rows"]] = case["dlogits_vals"]
    return d.view(B, T, V)


def _logits(case):
    return case["logits_i8"].float() * case["logits_scale"]


# ------------------------------------------------------------------------------- CPU: the oracle

@pytest.mark.parametrize("name", _case_ids())
def test_oracle_labels_and_focal_loss_equal_the_reference_run(name):
    from oracle.loss_oracle import focal_loss, mask_labels

    torch.set_num_threads(1)
    tok, cases = _cases()
    c = next(c for c in cases if c["name"] == name)
    labels = mask_labels(c["input_ids"], answer_token_id=tok["answer"],
                         endofchunk_token_id=tok["endofchunk"], media_token_id=tok["media"],
                         pad_token_id=tok["pad"])
    assert torch.equal(labels, c["labels"])
    z = _logits(c).requires_grad_(True)
    loss = focal_loss(z, labels, c["weights"], gamma=c["gamma"], use_reweight=c["use_reweight"])
    loss.backward()
    assert abs(float(loss.detach()) - float(c["loss"])) <= 1e-6 * abs(float(c["loss"]))
    assert rel_err(z.grad, _dense_grad(c)) < 1e-6
    # rows the reference gave no gradient stay exactly zero
    zero_rows = torch.ones(z.grad.flatten(0, 1).shape[0], dtype=torch.bool)
    zero_rows[c["dlogits_rows"]] = False
    assert z.grad.flatten(0, 1)[zero_rows].abs().max() == 0


def test_reference_unsqueezes_a_frame_dim_into_vision_x():
    # mmrec.py:135-137: patch_images (B,Ti,C,H,W) -> vision_x (B,Ti,1,C,H,W); our train.unimp_loss
    # does the same before calling the model
    _, cases = _cases()
    for c in cases:
        assert len(c["vision_x_shape"]) == 6 and c["vision_x_shape"][2] == 1


# ------------------------------------------------------------------------------- CPU: host rules

def test_collate_fn_equals_the_reference_collate():
    from unimp_b200.collate import collate_fn

    blob = torch.load(os.path.join(GOLDEN, "ref_collate.pt"), weights_only=False)
    got = collate_fn(blob["samples"], pad_idx=blob["pad_idx"], eos_idx=blob["eos_idx"])
    want = blob["batch"]
    assert set(got) == set(want) and set(got["net_input"]) == set(want["net_input"])
    for k, v in want["net_input"].items():
        assert got["net_input"][k].dtype == v.dtype, k
        assert torch.equal(got["net_input"][k], v), k
    assert collate_fn([], pad_idx=0, eos_idx=0) == {}


def test_weight_decay_groups_equal_the_reference_closure():
    from unimp_b200.train import apply_decay

    blob = torch.load(os.path.join(GOLDEN, "ref_host_rules.pt"), weights_only=False)
    assert any(blob["decay"].values()) and not all(blob["decay"].values())
    for name, want in blob["decay"].items():
        assert apply_decay(name) == want, name


def test_get_checkpoint_keys_equal_the_reference_function():
    from unimp_b200.train import get_checkpoint

    class _Toy(torch.nn.Module):   # same toy tree as make_reference_golden.py
        def __init__(self):
            super().__init__()
            self.frozen = torch.nn.Linear(3, 3)
            self.train_me = torch.nn.Linear(3, 2, bias=False)
            self.alias = torch.nn.ModuleList([self.frozen])
            self.register_buffer("buf", torch.zeros(2))
            for p in self.frozen.parameters():
                p.requires_grad_(False)

    blob = torch.load(os.path.join(GOLDEN, "ref_host_rules.pt"), weights_only=False)
    assert sorted(get_checkpoint(_Toy()).keys()) == blob["toy_checkpoint_keys"]
    # the documented fix for the alias quirk drops the frozen tensors under every name
    assert sorted(get_checkpoint(_Toy(), drop_frozen_aliases=True).keys()) == ["buf", "train_me.weight"]


# ------------------------------------------------------------------------------- GPU: the kernels

@pytest.mark.gpu
@pytest.mark.parametrize("name", _case_ids())
def test_cuda_mask_labels_equals_the_reference_run(name):
    from unimp_b200 import ops

    tok, cases = _cases()
    c = next(c for c in cases if c["name"] == name)
    got = ops.mask_labels(c["input_ids"].to(DEV), answer_token_id=tok["answer"],
                          endofchunk_token_id=tok["endofchunk"], media_token_id=tok["media"],
                          pad_token_id=tok["pad"])
    assert torch.equal(got.cpu(), c["labels"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol_l,tol_g", [(torch.float32, 1e-5, 1e-4), (torch.bfloat16, 1e-3, 2e-2)])
@pytest.mark.parametrize("name", _case_ids())
def test_cuda_focal_ce_equals_the_reference_run(name, dtype, tol_l, tol_g):
    from unimp_b200 import ops

    _, cases = _cases()
    c = next(c for c in cases if c["name"] == name)
    # the fixture's logits lie on a 1/8 grid: identical values in fp32 and bf16
    z = _logits(c).to(DEV, dtype).requires_grad_(True)
    loss = ops.focal_ce(z, c["labels"].to(DEV), c["weights"].to(DEV), gamma=c["gamma"],
                        use_focal=c["use_reweight"])
    loss.backward()
    assert abs(float(loss.detach()) - float(c["loss"])) <= tol_l * abs(float(c["loss"]))
    assert rel_err(z.grad, _dense_grad(c)) < tol_g
    zero_rows = torch.ones(z.grad.flatten(0, 1).shape[0], dtype=torch.bool)
    zero_rows[c["dlogits_rows"]] = False
    assert z.grad.flatten(0, 1)[zero_rows.to(DEV)].abs().max() == 0


# ------------------------------------------------------------------------------- the in-tree ViT tower

def _vit_fixture():
    blob = torch.load(os.path.join(GOLDEN, "ref_vit_tower.pt"), weights_only=False)
    ln = ("norm.weight", "norm1.weight", "norm2.weight")
    sd = {k: (1.0 if k.endswith(ln) else 0.0) + q.float() * blob["state_scale"]
          for k, q in blob["state_i8"].items()}
    return sd, blob["pixels_i8"].float() * blob["pixel_scale"], blob["last_hidden_state"]


def test_oracle_vision_tower_equals_the_reference_xformers_clip():
    """The oracle's tower (HF CLIPVisionModel) against reference `UniMP/xformers_model/clip.py`."""
    from transformers import CLIPVisionModel

    from unimp_b200 import tiny_config
    from util import hf_configs

    torch.set_num_threads(1)
    sd, pixels, want = _vit_fixture()
    vis = CLIPVisionModel(hf_configs(tiny_config())[0]).eval()
    missing, unexpected = vis.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    with torch.no_grad():
        got = vis(pixel_values=pixels).last_hidden_state
    assert rel_err(got, want) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_cuda_vision_tower_equals_the_reference_xformers_clip(dtype, tol):
    """Our ViT (tcgen05 attention for bf16, fused residual+LN, quick-GELU kernels) against the
    reference's in-tree tower; patch tokens = last_hidden_state[:, 1:] (no post-LN)."""
    from unimp_b200 import tiny_config
    from unimp_b200.vit import VisionTransformer, load_hf_clip_vision_weights

    cfg = tiny_config()
    sd, pixels, want = _vit_fixture()
    vit = VisionTransformer(image_size=cfg.image_size, patch_size=cfg.patch_size, width=cfg.vis_width,
                            layers=cfg.vis_layers, heads=cfg.vis_heads, mlp=cfg.vis_mlp).to(DEV, dtype)
    load_hf_clip_vision_weights(vit, {k: v.to(DEV, dtype) for k, v in sd.items()})
    _, tokens = vit(pixels.to(DEV, dtype))
    assert rel_err(tokens.float(), want[:, 1:]) < tol


# ------------------------------------------------------------------------------- real batch structure

def _dataset_fixture():
    return torch.load(os.path.join(GOLDEN, "ref_dataset_batches.pt"), weights_only=False)


def _classes(row, tok):
    """One letter per token: B bos/eos (same id in the GPT-NeoX vocabulary), M <image>, A <answer>,
    C <|endofchunk|>, P pad, i item / rating id, g image-token id, t plain text."""
    out = []
    for v in row.tolist():
        if v == tok["media"]:
            out.append("M")
        elif v == tok["answer"]:
            out.append("A")
        elif v == tok["endofchunk"]:
            out.append("C")
        elif v == tok["pad"]:
            out.append("P")
        elif v == tok["bos"]:
            out.append("B")
        elif tok["first_item"] <= v < tok["first_item"] + tok["n_items"]:
            out.append("i")
        elif tok["first_img"] <= v < tok["first_img"] + tok["n_img"]:
            out.append("g")
        else:
            out.append("t")
    return "".join(out)


REC_GRAMMAR = r"B(Mt+AiC)+t+AiBP*"        # rec_dataset.py:414,424,444-445 + collate right padding


@pytest.mark.parametrize("task", ["rec", "rate_exp", "img_gen"])
def test_oracle_labels_on_batches_built_by_the_reference_dataset_code(task):
    """Batches assembled by the reference's RecDataset sample builders + collate_fn; labels from the
    reference's train loop.  The oracle's state machine must agree bit for bit, and what survives
    the masking is exactly the answer spans plus the trailing EOS."""
    from oracle.loss_oracle import mask_labels

    blob = _dataset_fixture()
    tok, d = blob["tokens"], blob["tasks"][task]
    labels = mask_labels(d["input_ids"], answer_token_id=tok["answer"], endofchunk_token_id=tok["endofchunk"],
                         media_token_id=tok["media"], pad_token_id=tok["pad"])
    assert torch.equal(labels, d["labels"])
    assert d["weights"].tolist() == ([2.0] * 3 if task == "rec" else [1.0] * 3)   # rec_dataset.py:452 vs others
    for row, lab in zip(d["input_ids"], d["labels"]):
        cls = _classes(row, tok)
        kept = "".join(c for c, l in zip(cls, lab.tolist()) if l != -100)
        if task == "rec":
            import re
            assert re.fullmatch(REC_GRAMMAR, cls), cls
            assert kept == "i" * cls.count("A") + "B"          # one item per <answer>, then EOS
        elif task == "img_gen":
            # one retrieved image, one answer: the image-token ids ("img_N," - the reference joins them
            # with commas, rec_dataset.py:750-752) and the trailing EOS are what is scored
            assert cls.count("M") == 1 and cls.count("A") == 1
            assert kept == cls[cls.index("A") + 1:].rstrip("P") and kept.count("g") == 4 and kept.endswith("B")
        else:
            assert kept.endswith("B") and kept.count("i") == cls.count("A")   # rate id + explanation words


@pytest.mark.parametrize("name", ["C1-tiny", "C2-rec"])
def test_synthetic_batches_follow_the_grammar_of_the_reference_dataset(name):
    """unimp_b200/synth.py (what bench.py and the parity tests feed the model) produces rows of the
    same token grammar as the reference's rec sample builder, with the same label survivors."""
    import re

    from oracle.loss_oracle import mask_labels
    from unimp_b200 import openflamingo_4b_config, tiny_config
    from unimp_b200.config import WORKLOADS
    from unimp_b200.synth import make_batch

    cfg = tiny_config() if name == "C1-tiny" else openflamingo_4b_config()
    tk = cfg.tokens
    tok = {"answer": tk.answer, "endofchunk": tk.endofchunk, "media": tk.media, "pad": tk.pad, "bos": tk.bos,
           "first_item": tk.first_item, "n_items": tk.n_items, "first_img": tk.first_img, "n_img": tk.n_img}
    wl = WORKLOADS[name]
    b = make_batch(cfg, wl, seed=5)
    labels = mask_labels(b["input_ids"], answer_token_id=tk.answer, endofchunk_token_id=tk.endofchunk,
                         media_token_id=tk.media, pad_token_id=tk.pad)
    for row, lab in zip(b["input_ids"], labels):
        cls = _classes(row, tok)
        assert re.fullmatch(REC_GRAMMAR, cls), cls
        assert cls.count("M") == wl.Ti
        kept = "".join(c for c, l in zip(cls, lab.tolist()) if l != -100)
        assert kept == "i" * cls.count("A") + "B"
    assert b["weights"].tolist() == [2.0] * wl.B                  # the rec task weight


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["rec", "rate_exp", "img_gen"])
def test_cuda_mask_labels_on_batches_built_by_the_reference_dataset_code(task):
    from unimp_b200 import ops

    blob = _dataset_fixture()
    tok, d = blob["tokens"], blob["tasks"][task]
    got = ops.mask_labels(d["input_ids"].to(DEV), answer_token_id=tok["answer"], endofchunk_token_id=tok["endofchunk"],
                          media_token_id=tok["media"], pad_token_id=tok["pad"])
    assert torch.equal(got.cpu(), d["labels"])
