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
