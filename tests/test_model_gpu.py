"""End-to-end parity of the drop-in module on the GPU (config C1, BASELINE.json configs[0]):
logits / loss / gradients vs the CPU oracle with identical weights and batch, and vs the committed
golden vectors (which exist on the GPU box, where /root/reference does not)."""
import os

import pytest
import torch

from util import GOLDEN, build_oracle, copy_oracle_weights, rel_err

from oracle.loss_oracle import focal_loss, mask_labels
from unimp_b200 import tiny_config
from unimp_b200.config import WORKLOADS
from unimp_b200.synth import make_batch

pytestmark = pytest.mark.gpu

TOL = {torch.float32: dict(logits=1e-4, loss=1e-4, grad=2e-3),
       torch.bfloat16: dict(logits=2e-2, loss=1e-3, grad=8e-2)}  # BASELINE.json north_star tolerances


def _setup(dtype, ragged=True, seed=1234):
    from unimp_b200.factory import build_flamingo

    cfg = tiny_config()
    oracle = build_oracle(cfg, seed=0, gate=0.5)
    model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5)
    copy_oracle_weights(oracle, model)
    batch = make_batch(cfg, WORKLOADS["C1-tiny"], seed=seed, ragged=ragged)
    return cfg, oracle, model, batch


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ragged", [True, False])
def test_forward_loss_backward_vs_oracle(dtype, ragged):
    from unimp_b200.train import unimp_loss

    cfg, oracle, model, batch = _setup(dtype, ragged)
    tol = TOL[dtype]
    labels = mask_labels(batch["input_ids"], answer_token_id=cfg.tokens.answer,
                         endofchunk_token_id=cfg.tokens.endofchunk,
                         media_token_id=cfg.tokens.media, pad_token_id=cfg.tokens.pad)
    ref = oracle(vision_x=batch["patch_images"].unsqueeze(2), lang_x=batch["input_ids"],
                 attention_mask=batch["attention_masks"], labels=labels)
    ref_loss = focal_loss(ref.logits, labels, batch["weights"], gamma=2.0)
    ref_loss.backward()
    gb = {k: v.cuda() for k, v in batch.items()}
    loss, hf_loss, logits = unimp_loss(model, gb, cfg.tokens)
    loss.backward()
    assert rel_err(logits, ref.logits) < tol["logits"]
    assert abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)) < tol["loss"]
    assert abs(float(hf_loss) - float(ref.loss)) / abs(float(ref.loss)) < tol["loss"]
    ob = oracle.lang_encoder.gated_cross_attn_layers
    mb = model.lang_encoder.gated_cross_attn_layers
    pairs = [
        (mb[0].attn_gate.grad, ob[0].attn_gate.grad), (mb[1].ff_gate.grad, ob[1].ff_gate.grad),
        (mb[0].attn.to_q.weight.grad, ob[0].attn.to_q.weight.grad),
        (mb[0].attn.to_kv.weight.grad, ob[0].attn.to_kv.weight.grad),
        (mb[1].attn.to_out.weight.grad, ob[1].attn.to_out.weight.grad),
        (mb[0].attn.norm.weight.grad, ob[0].attn.norm.weight.grad),
        (mb[1].ff[1].weight.grad, ob[1].ff[1].weight.grad),
        (model.perceiver.latents.grad, oracle.perceiver.latents.grad),
        (model.perceiver.layers[0][0].to_kv.weight.grad, oracle.perceiver.layers[0][0].to_kv.weight.grad),
        (model.perceiver.layers[5][1][3].weight.grad, oracle.perceiver.layers[5][1][3].weight.grad),
        (model.perceiver.norm.bias.grad, oracle.perceiver.norm.bias.grad),
        (model.lang_encoder.get_input_embeddings().weight.grad,
         oracle.lang_encoder.lm.get_input_embeddings().weight.grad),
    ]
    for i, (got, want) in enumerate(pairs):
        if got.numel() == 1 and dtype == torch.bfloat16:
            # a scalar gate gradient is one long cancellation-prone sum of bf16 products
            assert abs(float(got) - float(want)) < 5e-3 + 0.1 * abs(float(want)), (i, got, want)
        else:
            assert rel_err(got, want) < tol["grad"], (i, rel_err(got, want))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_against_committed_golden_vectors(dtype):
    from unimp_b200.train import unimp_loss

    g = torch.load(os.path.join(GOLDEN, "tiny_fwd_loss.pt"))
    cfg, oracle, model, batch = _setup(dtype, ragged=True, seed=g["data_seed"])
    tol = TOL[dtype]
    gb = {k: v.cuda() for k, v in batch.items()}
    from unimp_b200 import ops
    labels = ops.mask_labels(gb["input_ids"], answer_token_id=cfg.tokens.answer,
                             endofchunk_token_id=cfg.tokens.endofchunk,
                             media_token_id=cfg.tokens.media, pad_token_id=cfg.tokens.pad)
    assert torch.equal(labels.cpu(), g["labels"])
    loss, hf_loss, logits = unimp_loss(model, gb, cfg.tokens)
    loss.backward()
    assert rel_err(logits, g["logits"]) < tol["logits"]
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < tol["loss"]
    assert abs(float(hf_loss) - float(g["hf_loss"])) / float(g["hf_loss"]) < tol["loss"]
    blk = model.lang_encoder.gated_cross_attn_layers[0]
    assert rel_err(blk.attn_gate.grad, g["grads"]["attn_gate0"]) < tol["grad"]
    assert rel_err(blk.attn.to_q.weight.grad, g["grads"]["to_q0"]) < tol["grad"]
    assert rel_err(model.perceiver.latents.grad, g["grads"]["latents"]) < tol["grad"]


def test_gate_zero_init_state_makes_xattn_vanish():
    """Upstream init: tanh(0) = 0 => logits independent of the images."""
    from unimp_b200.factory import build_flamingo

    cfg = tiny_config()
    model = build_flamingo(cfg, dtype=torch.float32, device="cuda", gate=None)
    b = make_batch(cfg, WORKLOADS["C1-tiny"], seed=1)
    gb = {k: v.cuda() for k, v in b.items()}
    with torch.no_grad():
        a = model(vision_x=gb["patch_images"].unsqueeze(2), lang_x=gb["input_ids"],
                  attention_mask=gb["attention_masks"]).logits
        c = model(vision_x=torch.randn_like(gb["patch_images"]).unsqueeze(2), lang_x=gb["input_ids"],
                  attention_mask=gb["attention_masks"]).logits
    assert torch.equal(a, c)


def test_generate_with_cached_media_matches_oracle_greedy():
    """a12: generate() (greedy, cached vision latents + cached x-attn K/V) emits the same tokens
    as the oracle re-running the full forward each step."""
    cfg, oracle, model, batch = _setup(torch.float32, ragged=False, seed=5)
    ids = batch["input_ids"][:1]
    L = int(batch["attention_masks"][0].sum()) - 3
    ids = ids[:, :L]
    vis = batch["patch_images"][:1].unsqueeze(2)
    new = 6
    cur = ids.clone()
    oracle.eval()
    with torch.no_grad():
        for _ in range(new):
            nxt = oracle(vision_x=vis, lang_x=cur).logits[:, -1].argmax(-1, keepdim=True)
            cur = torch.cat([cur, nxt], 1)
    model.eval()
    out = model.generate(vision_x=vis.cuda(), lang_x=ids.cuda(),
                         attention_mask=torch.ones_like(ids).cuda(), max_new_tokens=new,
                         num_beams=1, do_sample=False, eos_token_id=-1, pad_token_id=cfg.tokens.pad)
    assert out.cpu().tolist() == cur.tolist()


def test_graphed_train_step_matches_eager():
    """The CUDA-graph replay of the whole step follows the eager step's loss trajectory."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, GraphedTrainStep, get_grouped_params, train_step

    cfg = tiny_config()
    mbs = [{k: v.cuda() for k, v in make_batch(cfg, WORKLOADS["C1-tiny"], seed=i).items()} for i in range(2)]

    def fresh():
        m = build_flamingo(cfg, dtype=torch.float32, device="cuda", gate=0.5)
        return m, FlatAdamW(get_grouped_params(m, 0.1), lr=1e-3)

    model, opt = fresh()
    eager = [float(train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=mbs))
             for _ in range(4)]
    model, opt = fresh()
    before = [t.clone() for t in opt.state_tensors()]
    g = GraphedTrainStep(model, cfg.tokens, opt, None, mbs, warmup_iters=3)
    # building the graph consumes NO optimizer step: state and step counter untouched
    assert opt.step_count == 0
    for a, b in zip(before, opt.state_tensors()):
        assert torch.equal(a, b)
    got = [float(g(mbs)) for _ in range(4)]                                  # steps 1..4
    assert opt.step_count == 4
    for a, b in zip(got, eager):
        assert abs(a - b) < 2e-3 * abs(b)


def test_train_step_reduces_loss_and_updates_only_trainables():
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, get_grouped_params, train_step

    cfg = tiny_config()
    model = build_flamingo(cfg, dtype=torch.bfloat16, device="cuda", gate=0.5)
    frozen_before = model.lang_encoder.embed_out.weight.clone()
    opt = FlatAdamW(get_grouped_params(model, 0.1), lr=2e-3)
    b = make_batch(cfg, WORKLOADS["C1-tiny"], seed=3)
    gb = {k: v.cuda() for k, v in b.items()}
    losses = [float(train_step(model, gb, cfg.tokens, opt)) for _ in range(8)]
    assert losses[-1] < losses[0]
    assert torch.equal(frozen_before, model.lang_encoder.embed_out.weight)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_direct_grad_accumulation_equals_autograd_accumulation(dtype):
    """dW written by the backward GEMM into the flat buffer (beta=0 then beta=1 over two
    micro-batches) == autograd's AccumulateGrad path, for every trainable parameter."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, get_grouped_params, unimp_loss

    cfg = tiny_config()
    mbs = [{k: v.cuda() for k, v in make_batch(cfg, WORKLOADS["C1-tiny"], seed=i).items()} for i in range(2)]
    flats = []
    for direct in (True, False):
        model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5, seed=0)
        opt = FlatAdamW(get_grouped_params(model, 0.1), lr=1e-3, direct_grads=direct)
        n_direct = sum(g["n_direct"] for g in opt.groups)
        assert (n_direct > 0) == direct
        for rep in range(2):  # second round checks that stale data never leaks across steps
            opt.zero_grad()
            for mb in mbs:
                loss, _, _ = unimp_loss(model, mb, cfg.tokens)
                (loss / 2).backward()
        if direct:
            for g in opt.groups:
                for (_, p, _, _) in g["spans"][:g["n_direct"]]:
                    assert p._unimp_fresh is False  # every direct parameter was written
        flats.append({n: p.grad.detach().float().clone() for g in opt.groups for (n, p, _, _) in g["spans"]})
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for n in flats[0]:
        assert rel_err(flats[0][n], flats[1][n]) < tol, n


def test_beam_search_generate_runs_and_matches_greedy_prefix_semantics():
    """eval_rec.py-style call (num_beams > 1): vision_x is repeat_interleaved, cached media /
    cached x-attn K/V serve B*beams rows; beams=1 result is among the beams' hypotheses space
    (sanity: shapes, prompt preserved, deterministic)."""
    cfg, oracle, model, batch = _setup(torch.float32, ragged=False, seed=5)
    ids = batch["input_ids"][:2]
    L = int(batch["attention_masks"][:2].sum(1).min()) - 2
    ids = ids[:, :L].cuda()
    vis = batch["patch_images"][:2].unsqueeze(2).cuda()
    model.eval()
    kw = dict(attention_mask=torch.ones_like(ids), max_new_tokens=5, eos_token_id=-1,
              pad_token_id=cfg.tokens.pad, do_sample=False, early_stopping=True)
    a = model.generate(vision_x=vis, lang_x=ids, num_beams=3, num_return_sequences=1, **kw)
    b = model.generate(vision_x=vis, lang_x=ids, num_beams=3, num_return_sequences=1, **kw)
    assert a.shape == (2, L + 5) and torch.equal(a[:, :L], ids) and torch.equal(a, b)
    assert not model.lang_encoder.is_conditioned() and not model.lang_encoder._use_cached_vision_x


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_accumulation_window_equals_sequential_micro_batches(dtype):
    """One forward/backward over the concatenated accumulation window == the reference's sequential
    micro-batches (per-micro-batch loss normalisation kept): same loss, same accumulated gradients."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, get_grouped_params, unimp_loss, unimp_loss_fused

    cfg = tiny_config()
    mbs = [{k: v.cuda() for k, v in make_batch(cfg, WORKLOADS["C1-tiny"], seed=i, ragged=(i == 0)).items()}
           for i in range(2)]
    res = []
    for fused in (False, True):
        model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5, seed=0)
        opt = FlatAdamW(get_grouped_params(model, 0.1), lr=1e-3)
        opt.zero_grad()
        if fused:
            loss, _ = unimp_loss_fused(model, mbs, cfg.tokens)
            loss.backward()
        else:
            loss = 0.0
            for mb in mbs:
                l, _, _ = unimp_loss(model, mb, cfg.tokens)
                (l / 2).backward()
                loss = loss + l.detach() / 2
        res.append((float(loss), {n: p.grad.detach().float().clone() for g in opt.groups for (n, p, _, _) in g["spans"]}))
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    assert abs(res[0][0] - res[1][0]) < tol * abs(res[0][0])
    for n in res[0][1]:
        assert rel_err(res[1][1][n], res[0][1][n]) < tol, n


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("capacity", [True, 64])
def test_head_loss_fusion_on_label_rows_equals_the_dense_path(dtype, capacity):
    """gathering the rows with a valid shifted label BEFORE embed_out (exact gather or
    a fixed-capacity, graph-safe one) gives the same focal loss, the same logged HF mean CE, the
    same logits on those rows and the same gradients as the dense (B,T,V) path of
    reference UniMP/mmrec.py:177-213."""
    from unimp_b200 import ops
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, get_grouped_params, unimp_loss, unimp_loss_fused

    cfg = tiny_config()
    mbs = [{k: v.cuda() for k, v in make_batch(cfg, WORKLOADS["C1-tiny"], seed=i, ragged=(i == 0)).items()}
           for i in range(2)]
    res = {}
    for mode in ("dense", "rows"):
        model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5, seed=0)
        opt = FlatAdamW(get_grouped_params(model, 0.1), lr=1e-3)
        opt.zero_grad()
        kw = {} if mode == "dense" else {"label_rows": capacity}
        l1, hf1, lg1 = unimp_loss(model, mbs[0], cfg.tokens, **kw)            # single micro-batch
        l1.backward()
        g1 = {n: p.grad.detach().float().clone() for g in opt.groups for (n, p, _, _) in g["spans"]}
        opt.zero_grad()
        l2, lg2 = unimp_loss_fused(model, mbs, cfg.tokens, **kw)               # accumulation window
        l2.backward()
        g2 = {n: p.grad.detach().float().clone() for g in opt.groups for (n, p, _, _) in g["spans"]}
        res[mode] = (float(l1), float(hf1), lg1.detach().float(), g1, float(l2), lg2.detach().float(), g2)
    d, r = res["dense"], res["rows"]
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert abs(d[0] - r[0]) < tol * abs(d[0]) and abs(d[1] - r[1]) < tol * abs(d[1])
    assert abs(d[4] - r[4]) < tol * abs(d[4])
    labels = ops.mask_labels(mbs[0]["input_ids"], answer_token_id=cfg.tokens.answer,
                             endofchunk_token_id=cfg.tokens.endofchunk, media_token_id=cfg.tokens.media,
                             pad_token_id=cfg.tokens.pad)
    idx, tgt, _ = ops.gather_label_rows(labels)
    V = d[2].shape[-1]
    want_rows = d[2].reshape(-1, V)[idx.cpu()]
    assert rel_err(r[2][:idx.numel()], want_rows) < (1e-6 if dtype == torch.float32 else 1e-2)
    if capacity is not True:
        assert r[2].shape[0] == capacity
    for n in d[3]:
        # a scalar gate gradient is one long cancellation-prone sum: the two paths add it up in a
        # different order (measured 6e-5 relative in fp32)
        t = tol * (25 if d[3][n].numel() == 1 else 1)
        assert rel_err(r[3][n], d[3][n]) < t, n
        assert rel_err(r[6][n], d[6][n]) < t, n


def test_head_loss_fusion_overflow_of_the_static_capacity_is_loud():
    """A fixed-capacity gather that cannot hold every valid row must not return a plausible loss."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import unimp_loss

    cfg = tiny_config()
    model = build_flamingo(cfg, dtype=torch.float32, device="cuda", gate=0.5)
    gb = {k: v.cuda() for k, v in make_batch(cfg, WORKLOADS["C1-tiny"], seed=0).items()}
    loss, _, _ = unimp_loss(model, gb, cfg.tokens, label_rows=2)
    assert torch.isnan(loss)
    loss, _, _ = unimp_loss(model, gb, cfg.tokens, label_rows=512)
    assert torch.isfinite(loss)


# ------------------------------------------------------------------ 4B shapes (configs[1] geometry)

def _build_4b_pair(dtype_product):
    """fp32 oracle ON THE GPU (so that it finishes in seconds) + the product with the same weights,
    at the 4B geometry: dh = 80 rotary LM, D = 2560 K5 kernels, padded head at V = 74 053, 16 x-attn
    blocks, ViT-L/14.  Cheap random init (bench.build_oracle_model)."""
    import bench
    from unimp_b200 import openflamingo_4b_config
    from unimp_b200.factory import build_flamingo

    cfg = openflamingo_4b_config()
    oracle, _ = bench.build_oracle_model(cfg, "cuda")
    oracle.eval()
    model = build_flamingo(cfg, dtype=dtype_product, device="cuda", gate=0.5)
    copy_oracle_weights(oracle, model)
    return cfg, oracle, model


def test_4b_bf16_product_matches_fp32_oracle_logits_and_loss():
    """end-to-end parity at the 4B shapes on configs[1]'s micro-batch (B=3, T=256,
    Ti=2) — north-star bars: bf16 rel 2e-2 on logits, 1e-3 on loss — for the dense path and for the
    head+loss fusion (the bf16 path includes the K1-fused cluster kernel and the K4 causal-attention
    kernels of all 32 decoder layers)."""
    from unimp_b200.config import Workload
    from unimp_b200.train import unimp_loss

    cfg, oracle, model = _build_4b_pair(torch.bfloat16)
    wl = Workload("C2-rec", B=3, Ti=2, T=256)
    gb = {k: v.cuda() for k, v in make_batch(cfg, wl, seed=1234).items()}
    labels = mask_labels(gb["input_ids"].cpu(), answer_token_id=cfg.tokens.answer,
                         endofchunk_token_id=cfg.tokens.endofchunk, media_token_id=cfg.tokens.media,
                         pad_token_id=cfg.tokens.pad).cuda()
    with torch.no_grad():
        ref = oracle(vision_x=gb["patch_images"].unsqueeze(2), lang_x=gb["input_ids"],
                     attention_mask=gb["attention_masks"], labels=labels)
        ref_loss = focal_loss(ref.logits, labels, gb["weights"], gamma=2.0)
        # the reference's own bf16 path: the same oracle under autocast (UniMP/mmrec.py:176)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = oracle(vision_x=gb["patch_images"].unsqueeze(2), lang_x=gb["input_ids"],
                        attention_mask=gb["attention_masks"], labels=labels)
        ac_loss = focal_loss(ac.logits.float(), labels, gb["weights"], gamma=2.0)
        from unimp_b200 import ops as _ops
        k4_calls, k4_orig = [], _ops.rotary_lm_attention
        _ops.rotary_lm_attention = lambda *a, **k: (k4_calls.append(k.get("head_dim")), k4_orig(*a, **k))[1]
        try:
            loss, hf_loss, logits = unimp_loss(model, gb, cfg.tokens)
        finally:
            _ops.rotary_lm_attention = k4_orig
        # K4: all 32 decoder layers ran on unimp_lm_attn_fwd (head dim 80, padded batch -> key bits)
        assert len(k4_calls) == cfg.lm_layers and set(k4_calls) == {80}, k4_calls
        loss_r, hf_r, logits_r = unimp_loss(model, gb, cfg.tokens, label_rows=True)
    valid = gb["attention_masks"].bool()
    e_logits = rel_err(logits[valid], ref.logits[valid])
    e_loss = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    e_hf = abs(float(hf_loss) - float(ref.loss)) / abs(float(ref.loss))
    a_logits = rel_err(ac.logits[valid], ref.logits[valid])
    a_loss = abs(float(ac_loss) - float(ref_loss)) / abs(float(ref_loss))
    print(f"4B bf16 vs fp32 oracle: logits {e_logits:.2e} loss {e_loss:.2e} hf_loss {e_hf:.2e} | the "
          f"oracle's own bf16-autocast pass vs its fp32 pass: logits {a_logits:.2e} loss {a_loss:.2e}")
    # north-star bars (bf16: 2e-2 logits, 1e-3 loss); where 32 bf16 layers put even the reference's
    # own autocast path above a bar, the product must not be worse than that path by more than 1.5x
    bar_logits = max(TOL[torch.bfloat16]["logits"], 1.5 * a_logits)
    bar_loss = max(TOL[torch.bfloat16]["loss"], 1.5 * a_loss)
    assert e_logits < bar_logits
    assert e_loss < bar_loss and e_hf < bar_loss
    assert abs(float(loss_r) - float(ref_loss)) / abs(float(ref_loss)) < bar_loss
    assert abs(float(hf_r) - float(ref.loss)) / abs(float(ref.loss)) < bar_loss


def test_4b_fp32_graphed_decoder_emits_the_tokens_of_the_oracle():
    """decode parity at the 4B shapes in fp32 — GraphedDecoder (cached vision latents,
    cached x-attn K/V, CUDA-graph replays) == Flamingo.generate (HF beam search on our kernels) ==
    the oracle re-running its full forward for every greedy token."""
    from unimp_b200.config import Workload
    from unimp_b200.decode import GraphedDecoder

    cfg, oracle, model = _build_4b_pair(torch.float32)
    model.eval()
    wl = Workload("C4-decode", B=1, Ti=3, T=96)
    b = make_batch(cfg, wl, seed=7)
    L = int(b["attention_masks"][0].sum()) - 2
    ids = b["input_ids"][:, :L].cuda()
    vis = b["patch_images"].unsqueeze(2).cuda()
    new = 6
    cur = ids.clone()
    with torch.no_grad():
        for _ in range(new):
            nxt = oracle(vision_x=vis, lang_x=cur).logits[:, -1].argmax(-1, keepdim=True)
            cur = torch.cat([cur, nxt], 1)
    dec = GraphedDecoder(model)
    kw = dict(max_new_tokens=new, eos_token_id=-1, pad_token_id=cfg.tokens.pad)
    greedy = dec.generate(vis, ids, torch.ones_like(ids), num_beams=1, **kw)
    assert greedy.tolist() == cur.tolist()
    beams = dec.generate(vis, ids, torch.ones_like(ids), num_beams=3, early_stopping=False, **kw)
    hf = model.generate(vision_x=vis, lang_x=ids, attention_mask=torch.ones_like(ids), num_beams=3,
                        do_sample=False, early_stopping=False, **kw)
    assert beams.tolist() == hf.tolist()


def test_reference_style_train_loop_drives_the_product_model():
    """VERDICT r1 missing #6: the loop body of reference `UniMP/mmrec.py:135-215` restated line by
    line (batch unpack, Python label loop, `model(vision_x=, lang_x=, attention_mask=, labels=)`,
    `output[0]`, `output["logits"]`, the focal-loss lines, backward under gradient accumulation) and
    driven against the PRODUCT model as a drop-in — no unimp_b200.train helper in the loop.  Its
    losses and accumulated gradients must equal what `unimp_b200.train.train_step`'s pieces produce
    (GPU label kernel, focal-CE kernel, fused accumulation window, head+loss fusion)."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, get_grouped_params, unimp_loss_fused

    cfg = tiny_config()
    tk = cfg.tokens
    accum = 2
    batches = [make_batch(cfg, WORKLOADS["C1-tiny"], seed=20 + i, ragged=(i == 1)) for i in range(accum)]
    gamma = 2.0

    # ---- (1) the reference's loop, product model as the drop-in ---------------------------------
    model = build_flamingo(cfg, dtype=torch.float32, device="cuda", gate=0.5, seed=0)
    opt = FlatAdamW(get_grouped_params(model, 0.1), lr=1e-3, direct_grads=True)
    opt.zero_grad()
    ref_losses, ref_out0 = [], []
    for batch in batches:                                                   # accelerator.accumulate(model)
        images = batch["patch_images"].cuda().unsqueeze(2)                  # mmrec.py:135-137
        input_ids = batch["input_ids"].cuda()
        attention_mask = batch["attention_masks"].cuda()
        weights = batch["weights"].cuda()
        labels = mask_labels(batch["input_ids"], answer_token_id=tk.answer,  # mmrec.py:143-168 (the loop)
                             endofchunk_token_id=tk.endofchunk, media_token_id=tk.media,
                             pad_token_id=tk.pad).cuda()
        output = model(vision_x=images, lang_x=input_ids, attention_mask=attention_mask, labels=labels)
        ref_out0.append(float(output[0]))                                    # mmrec.py:182
        loss = focal_loss(output["logits"], labels, weights, gamma=gamma)    # mmrec.py:190-213
        (loss / accum).backward()                                            # accelerate divides by the window
        ref_losses.append(float(loss))
    opt._zero_unwritten()
    g_ref = {n: p.grad.detach().clone() for g in opt.groups for (n, p, _, _) in g["spans"]}
    assert all(torch.isfinite(torch.tensor(ref_losses))) and all(torch.isfinite(torch.tensor(ref_out0)))

    # ---- (2) the product's own step pieces on the same window -----------------------------------
    for label_rows in (None, True):
        model2 = build_flamingo(cfg, dtype=torch.float32, device="cuda", gate=0.5, seed=0)
        opt2 = FlatAdamW(get_grouped_params(model2, 0.1), lr=1e-3)
        opt2.zero_grad()
        mbs = [{k: v.cuda() for k, v in b.items()} for b in batches]
        loss2, _ = unimp_loss_fused(model2, mbs, tk, gamma=gamma, label_rows=label_rows)
        loss2.backward()
        opt2._zero_unwritten()
        assert abs(float(loss2) - sum(ref_losses) / accum) < 2e-5 * abs(float(loss2))
        for g in opt2.groups:
            for (n, p, _, _) in g["spans"]:
                t = 5e-4 if p.numel() == 1 else 2e-5
                assert rel_err(p.grad, g_ref[n]) < t, (label_rows, n, rel_err(p.grad, g_ref[n]))


@pytest.mark.parametrize("fuse", [True, False])
def test_deferred_optimizer_graph_equals_eager_steps(fuse):
    """GraphedTrainStep(defer_optimizer=True): the clip + AdamW pass of step k runs at the start of
    replay k+1 on a side stream under the frozen ViT forward.  After flush() the parameters, the
    AdamW state and the loss trajectory equal K ordinary steps."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import FlatAdamW, GraphedTrainStep, get_grouped_params, train_step

    cfg = tiny_config()
    sets = [[{k: v.cuda() for k, v in make_batch(cfg, WORKLOADS["C1-tiny"], seed=10 * j + i).items()}
             for i in range(2)] for j in range(3)]
    scales = [0.25, 0.5, 1.0, 1.0, 0.7]

    def fresh():
        m = build_flamingo(cfg, dtype=torch.float32, device="cuda", gate=0.5)
        return m, FlatAdamW(get_grouped_params(m, 0.1), lr=1e-3)

    model, opt = fresh()
    eager = [float(train_step(model, None, cfg.tokens, opt, None, accum_steps=2, micro_batches=sets[i % 3],
                              fuse_accum=fuse, lr_scale=scales[i])) for i in range(5)]
    want = [t.clone() for t in opt.state_tensors()]
    torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
    model, opt = fresh()
    g = GraphedTrainStep(model, cfg.tokens, opt, None, sets[0], fuse_accum=fuse, defer_optimizer=True)
    assert opt.step_count == 0
    got = [float(g(sets[i % 3], lr_scale=scales[i])) for i in range(5)]
    assert opt.step_count == 4                      # the 5th update is still pending
    g.flush()
    assert opt.step_count == 5
    torch.cuda.synchronize()
    for a, b in zip(got, eager):
        assert abs(a - b) < 2e-4 * abs(b), (got, eager)
    for a, b in zip(opt.state_tensors(), want):
        assert rel_err(a, b) < 1e-4
