"""Oracle self-consistency and its pinning to the committed golden vectors.

The reference has no tests or fixtures (parity unpinned); these checks are what anchors the
restatement: each one proves two independent formulations of the same arithmetic agree."""
import os

import pytest
import torch

from util import GOLDEN, build_oracle

from oracle.flamingo_oracle import MaskedCrossAttention
from oracle.loss_oracle import focal_loss, focal_loss_closed_form_grad, mask_labels
from unimp_b200 import tiny_config
from unimp_b200.config import WORKLOADS
from unimp_b200.synth import make_batch


def test_dense_masked_equals_block_local():
    """(i) dense masked einsum form == per-image local attention."""
    torch.manual_seed(0)
    B, T, D, Dv, Ti, n = 2, 24, 32, 16, 3, 64
    m = MaskedCrossAttention(dim=D, dim_visual=Dv).double()
    x = torch.randn(B, T, D, dtype=torch.float64)
    media = torch.randn(B, Ti, n, Dv, dtype=torch.float64)
    loc = torch.zeros(B, T, dtype=torch.bool)
    loc[0, [2, 9, 15]] = True
    loc[1, [5, 6]] = True  # sample 1: text before first image, only 2 of 3 images referenced
    dense = m(x, media, media_locations=loc)
    # local form
    xl = m.norm(x)
    q = m.to_q(xl).view(B, T, 8, 64) * m.scale
    kv = m.to_kv(media.view(B, Ti * n, Dv))
    k, v = kv.chunk(2, -1)
    k = k.view(B, Ti, n, 8, 64)
    v = v.view(B, Ti, n, 8, 64)
    tt = loc.cumsum(-1)
    out = torch.zeros(B, T, 8, 64, dtype=torch.float64)
    for b in range(B):
        for t in range(T):
            j = int(tt[b, t])
            if j == 0:
                continue
            s = torch.einsum("hd,nhd->hn", q[b, t], k[b, j - 1])
            out[b, t] = torch.einsum("hn,nhd->hd", s.softmax(-1), v[b, j - 1])
    local = m.to_out(out.view(B, T, 512))
    assert torch.allclose(dense, local, atol=1e-10)
    assert dense[1, :5].abs().max() == 0  # text before the first <image> gets exactly zero


def test_text_time_beyond_images_is_uniform():
    """Upstream quirk kept: more <image> tokens than images -> uniform attention over all keys."""
    torch.manual_seed(0)
    m = MaskedCrossAttention(dim=16, dim_visual=8).double()
    x = torch.randn(1, 4, 16, dtype=torch.float64)
    media = torch.randn(1, 1, 64, 8, dtype=torch.float64)
    loc = torch.tensor([[True, False, True, False]])
    out = m(x, media, media_locations=loc)
    v = m.to_kv(media.view(1, 64, 8)).chunk(2, -1)[1]
    uniform = m.to_out(v.mean(1, keepdim=True))
    assert torch.allclose(out[:, 2:], uniform.expand(1, 2, 16), atol=1e-10)


@pytest.mark.parametrize("gamma,use", [(2.0, True), (0.5, True), (2.0, False)])
def test_focal_closed_form_gradient_matches_autograd(gamma, use):
    """(ii)"""
    torch.manual_seed(1)
    B, T, V = 3, 9, 130
    z = torch.randn(B, T, V, dtype=torch.float64, requires_grad=True)
    y = torch.randint(0, V, (B, T))
    y[0, :4] = -100
    y[2, 5:] = -100
    w = torch.tensor([2.0, 1.0, 1.0], dtype=torch.float64)
    loss = focal_loss(z, y, w, gamma=gamma, use_reweight=use)
    (g,) = torch.autograd.grad(loss, z)
    g2 = focal_loss_closed_form_grad(z.detach(), y, w, gamma=gamma, use_reweight=use)
    assert torch.allclose(g, g2, atol=1e-12)


def test_gamma0_weight1_is_plain_ce_and_matches_hf_loss():
    """(iii) + (v)"""
    torch.manual_seed(2)
    B, T, V = 2, 7, 128
    z = torch.randn(B, T, V, dtype=torch.float64)
    y = torch.randint(0, V, (B, T))
    y[:, 0] = -100
    y[1, 4:] = -100
    ones = torch.ones(B, dtype=torch.float64)
    a = focal_loss(z, y, ones, gamma=0.0, use_reweight=True)
    b = torch.nn.functional.cross_entropy(z[:, :-1].reshape(-1, V), y[:, 1:].reshape(-1), ignore_index=-100)
    assert torch.allclose(a, b, atol=1e-12)


def test_mask_labels_state_machine_examples():
    A, E, M, P = 90, 91, 92, 93
    ids = torch.tensor([[1, M, 5, 6, A, 7, E, 8, M, 9, A, 10, 11, 0, P, P],
                        [1, 2, 3, A, A, 4, E, E, 5, A, M, 6, 0, P, P, P]])
    lab = mask_labels(ids, answer_token_id=A, endofchunk_token_id=E, media_token_id=M, pad_token_id=P)
    exp0 = [-100] * 5 + [7] + [-100] * 5 + [10, 11, 0, -100, -100]
    exp1 = [-100] * 5 + [4] + [-100] * 5 + [6, 0, -100, -100, -100]
    assert lab[0].tolist() == exp0
    assert lab[1].tolist() == exp1


def test_decode_with_cached_media_equals_full_forward():
    """(iv) cached-media single-token step == last-token logits of a full re-forward."""
    cfg = tiny_config()
    model = build_oracle(cfg).eval()
    batch = make_batch(cfg, WORKLOADS["C1-tiny"], seed=7)
    ids, am = batch["input_ids"], batch["attention_masks"]
    vis = batch["patch_images"].unsqueeze(2)
    L = int(am[0].sum())  # use sample 0 up to its true length, no padding
    ids0, vis0 = ids[:1, :L], vis[:1]
    with torch.no_grad():
        full = model(vision_x=vis0, lang_x=ids0).logits[:, -1]
        model.cache_media(ids0[:, :-1], vis0)
        out = model(vision_x=None, lang_x=ids0[:, :-1], use_cache=True, clear_conditioned_layers=False)
        step = model(vision_x=None, lang_x=ids0[:, -1:], past_key_values=out.past_key_values,
                     use_cache=True, clear_conditioned_layers=False).logits[:, -1]
        model.uncache_media()
    assert torch.allclose(full, step, atol=2e-5, rtol=1e-4)


def test_oracle_matches_committed_golden_vectors():
    """The committed fixtures were produced by tests/golden/make_golden.py from this oracle;
    this pins the oracle against silent drift (a transformers upgrade, an edit)."""
    path = os.path.join(GOLDEN, "tiny_fwd_loss.pt")
    g = torch.load(path)
    cfg = tiny_config()
    model = build_oracle(cfg, seed=g["weight_seed"], gate=g["gate"])
    batch = make_batch(cfg, WORKLOADS["C1-tiny"], seed=g["data_seed"], ragged=True)
    assert torch.equal(batch["input_ids"], g["input_ids"])
    labels = mask_labels(batch["input_ids"], answer_token_id=cfg.tokens.answer,
                         endofchunk_token_id=cfg.tokens.endofchunk,
                         media_token_id=cfg.tokens.media, pad_token_id=cfg.tokens.pad)
    assert torch.equal(labels, g["labels"])
    out = model(vision_x=batch["patch_images"].unsqueeze(2), lang_x=batch["input_ids"],
                attention_mask=batch["attention_masks"], labels=labels)
    loss = focal_loss(out.logits, labels, batch["weights"], gamma=2.0)
    assert torch.allclose(out.logits, g["logits"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(loss, g["loss"], rtol=1e-5)
    assert torch.allclose(out.loss, g["hf_loss"], rtol=1e-5)
