"""Host-side logic that needs no GPU: synthetic batch layout, weight-decay rule, LR schedule, and the
N>1 data-parallel path (bucketed all-reduce) with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.loss_oracle import mask_labels
from unimp_b200 import openflamingo_4b_config, tiny_config
from unimp_b200.config import WORKLOADS
from unimp_b200.synth import make_batch
from unimp_b200.train import BucketedAllReduce, FlatAdamW, apply_decay, cosine_with_warmup


def test_vocab_layout_matches_reference_token_growth():
    cfg = openflamingo_4b_config()
    t = cfg.tokens
    # reference UniMP/mmrec.py:538-581: <answer>, 5 rate_*, 5 s_*, 22738 item_*, 1024 img_*
    assert cfg.vocab == 50277 + 3 + 1 + 5 + 5 + 22738 + 1024 == 74053
    assert t.first_item - t.answer - 1 == 10 and t.first_img == t.first_item + 22738


def test_vocab_layout_equals_the_reference_vocab_growth_run():
    """`tests/golden/ref_host_rules.pt::vocab_growth` = the tokens `main()` appends to the tokenizer,
    obtained by executing `UniMP/mmrec.py:538-581` unmodified against a recording tokenizer.  Ids
    are positional (upstream's factory adds <|endofchunk|>, <image>, <PAD> first), so the order of
    that list IS the id layout `openflamingo_4b_config()` must encode."""
    import hashlib
    import os

    from util import GOLDEN

    v = torch.load(os.path.join(GOLDEN, "ref_host_rules.pt"), weights_only=False)["vocab_growth"]
    cfg = openflamingo_4b_config()
    t = cfg.tokens
    base = 50277                                  # RedPajama-INCITE tokenizer length
    assert (t.endofchunk, t.media, t.pad) == (base, base + 1, base + 2)
    first_added = base + 3
    starts = {g[0]: (first_added + g[1], g[2]) for g in v["groups"]}
    assert starts["<answer>"] == (t.answer, 1)
    assert starts["item_"] == (t.first_item, t.n_items)
    assert starts["img_"] == (t.first_img, t.n_img)
    assert first_added + v["n_added"] == cfg.vocab == 74053
    # the exact token strings, in order (what a real tokenizer would be given)
    want = (["<answer>"] + [f"rate_{i}" for i in range(1, 6)] + [f"s_{i}" for i in range(5)]
            + [f"item_{i}" for i in range(t.n_items)] + [f"img_{i}," for i in range(t.n_img)])
    assert hashlib.sha256("\n".join(want).encode()).hexdigest() == v["sha256"]


@pytest.mark.parametrize("name", ["C1-tiny", "C2-rec", "C5-imggen"])
def test_synthetic_batch_has_the_reference_prompt_structure(name):
    cfg = tiny_config() if name == "C1-tiny" else openflamingo_4b_config()
    wl = WORKLOADS[name]
    b = make_batch(cfg, wl, seed=3, ragged=True)
    ids, am = b["input_ids"], b["attention_masks"]
    assert ids.shape == (wl.B, wl.T) and b["patch_images"].shape[:2] == (wl.B, wl.Ti)
    t = cfg.tokens
    for i in range(wl.B):
        L = int(am[i].sum())
        assert (ids[i, L:] == t.pad).all() and (am[i, :L] == 1).all()      # right padding
        n_img = int((ids[i] == t.media).sum())
        assert n_img == (max(1, wl.Ti - 1) if i == 0 else wl.Ti)           # ragged sample 0
        assert int((ids[i] == t.answer).sum()) == n_img + 1                # one answer per chunk + final
        assert int((ids[i] == t.endofchunk).sum()) == n_img
    lab = mask_labels(ids, answer_token_id=t.answer, endofchunk_token_id=t.endofchunk,
                      media_token_id=t.media, pad_token_id=t.pad)
    valid = lab != -100
    assert valid.any()
    kept = ids[valid]
    lo, hi = t.first_item, t.first_img + t.n_img
    assert (((kept >= lo) & (kept < hi)) | (kept == t.eos)).all()          # answers (+ trailing EOS) only


def test_weight_decay_rule_is_the_reference_rule():
    # reference UniMP/mmrec.py:612-619
    assert apply_decay("lang_encoder.gated_cross_attn_layers.3.attn.to_q.weight")
    assert apply_decay("lang_encoder.gpt_neox.layers.1.gated_cross_attn_layer.ff.1.weight")
    assert not apply_decay("lang_encoder.gated_cross_attn_layers.3.attn_gate")
    assert not apply_decay("lang_encoder.gated_cross_attn_layers.3.ff_gate")
    assert not apply_decay("lang_encoder.gated_cross_attn_layers.3.attn.norm.weight")
    assert not apply_decay("perceiver.layers.0.0.to_q.weight")
    assert not apply_decay("lang_encoder.gpt_neox.embed_in.weight")


def test_cosine_schedule_matches_transformers():
    from transformers import get_cosine_schedule_with_warmup

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = get_cosine_schedule_with_warmup(opt, num_warmup_steps=5, num_training_steps=40)
    for s in range(40):
        assert abs(opt.param_groups[0]["lr"] - cosine_with_warmup(s, 5, 40)) < 1e-7
        opt.step()
        sch.step()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, accum, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical weights on every rank (bench: weights seed 0)
    net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 8))
    groups = [{"params": [(n, p) for n, p in net.named_parameters() if "weight" in n], "weight_decay": 0.1},
              {"params": [(n, p) for n, p in net.named_parameters() if "bias" in n], "weight_decay": 0.0}]
    opt = FlatAdamW(groups, lr=1e-3)
    red = BucketedAllReduce(opt, bucket_bytes=1024)  # several buckets
    assert len(red.buckets) >= 3
    opt.zero_grad()
    g = torch.Generator().manual_seed(100 + rank)   # per-rank data (reference: seed + rank)
    xs = [torch.randn(5, 24, generator=g) for _ in range(accum)]
    for i, x in enumerate(xs):
        red.armed = i == accum - 1
        (net(x).pow(2).mean() / accum).backward()
    red.finish()
    flat = torch.cat([grp["flat_g"] for grp in opt.groups]) * red.grad_scale
    q.put((rank, flat.clone(), [x.clone() for x in xs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("accum", [1, 2])
def test_bucketed_allreduce_world2_equals_single_process_mean(accum):
    """N-GPU == 1-GPU gradient equality on the same global batch, over gloo."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, accum, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.allclose(res[0][1], res[1][1], atol=0, rtol=0)  # every rank holds the same mean
    # single-process reference: average of the per-rank losses (DeepSpeed DP semantics)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 8))
    groups = [{"params": [(n, p) for n, p in net.named_parameters() if "weight" in n], "weight_decay": 0.1},
              {"params": [(n, p) for n, p in net.named_parameters() if "bias" in n], "weight_decay": 0.0}]
    opt = FlatAdamW(groups, lr=1e-3)
    opt.zero_grad()
    for r in range(world):
        for x in res[r][2]:
            (net(x).pow(2).mean() / accum / world).backward()
    want = torch.cat([grp["flat_g"] for grp in opt.groups])
    assert torch.allclose(res[0][1], want, atol=1e-7, rtol=1e-5)


def _cpu_optimizer_kernels():
    """Test-side stand-ins for `unimp_sumsq` / `unimp_adamw_step` (the arithmetic of
    unimp_b200/csrc/misc.cu::adamw_kernel, restated with torch) so that the COLLECTIVE plumbing of
    the sharded optimizer can run under gloo on a GPU-less box.  The kernels themselves are tested
    against torch.optim.AdamW on the GPU (tests/test_kernels_gpu.py::test_fused_adamw_matches_torch)."""
    from unimp_b200 import ops

    def sumsq_(grad, acc):
        acc += grad.float().pow(2).sum()

    def adamw_step_(master, param, grad, m, v, *, hyper, beta1, beta2, eps, weight_decay, gnorm_sq=None,
                    max_norm=0.0, grad_scale=1.0, background=False):
        lr, bc1, bc2_sqrt = (float(x) for x in hyper)
        clip = grad_scale
        if gnorm_sq is not None:
            clip *= min(1.0, max_norm / (float(gnorm_sq.sqrt()) * grad_scale + 1e-6))
        g = grad.float() * clip
        master.mul_(1.0 - lr * weight_decay)
        m.mul_(beta1).add_(g, alpha=1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        master.sub_((lr / bc1) * (m / (v.sqrt() / bc2_sqrt + eps)))
        param.copy_(master)

    ops.sumsq_, ops.adamw_step_ = sumsq_, adamw_step_


def _toy_net_and_groups():
    torch.manual_seed(0)  # identical weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 8))
    groups = [{"params": [(n, p) for n, p in net.named_parameters() if "weight" in n], "weight_decay": 0.1},
              {"params": [(n, p) for n, p in net.named_parameters() if "bias" in n], "weight_decay": 0.0}]
    return net, groups


def _sharded_worker(rank, world, port, steps, deferred, q):
    from unimp_b200.train import ShardedDataParallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _cpu_optimizer_kernels()
    net, groups = _toy_net_and_groups()
    opt = FlatAdamW(groups, lr=1e-2, shard_world=world, allocate_states=False)
    red = ShardedDataParallel(opt, bucket_bytes=1024, deferred_gather_module=net[0] if deferred else None)
    assert len(red.buckets) >= 3 and len(red.shards) == len(red.buckets)
    for sh in red.shards:   # a rank owns exactly 1/world of every bucket, 16-byte aligned
        assert sh["g_shard"].numel() * world == sh["g_full"].numel()
        assert sh["g_shard"].data_ptr() % 16 == 0 and sh["master"].numel() == sh["g_shard"].numel()
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(5, 24, generator=g) for _ in range(steps)]
    for x in xs:
        red.begin_step()
        opt.zero_grad()
        red.armed = True
        net(x).pow(2).mean().backward()
        opt.step_with(red, lr_scale=1.0)
    red.sync_params()
    flat = torch.cat([grp["flat_p"] for grp in opt.groups])
    q.put((rank, flat.clone(), [x.clone() for x in xs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("deferred", [False, True])
def test_sharded_data_parallel_world2_equals_single_process_adamw(deferred):
    """The default N>1 path (reduce-scatter, AdamW on the 1/N shard, all-gather — immediate or
    deferred to the next step's forward) over gloo with world_size 2: after 3 steps every rank
    holds the parameters a single process gets from AdamW on the rank-averaged gradients."""
    world, steps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, steps, deferred, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0][1], res[1][1])            # replicas stay bit-identical
    # single-process reference with the same stand-in kernels: mean of the per-rank losses
    _cpu_optimizer_kernels()
    net, groups = _toy_net_and_groups()
    opt = FlatAdamW(groups, lr=1e-2, shard_world=world)
    for s in range(steps):
        opt.zero_grad()
        for r in range(world):
            (net(res[r][2][s]).pow(2).mean() / world).backward()
        opt.step()
    want = torch.cat([grp["flat_p"] for grp in opt.groups])
    assert torch.allclose(res[0][1], want, atol=1e-6, rtol=1e-5)
    assert (res[0][1] - want).abs().max() < 1e-5 and (want != 0).any()


def test_get_checkpoint_keeps_trainables_with_upstream_key_names_and_reloads():
    """reference train_utils.py:258-265 + mmrec.py:513-514 (`load_state_dict(..., strict=False)`)."""
    from unimp_b200.factory import build_flamingo
    from unimp_b200.train import get_checkpoint

    cfg = tiny_config()
    m = build_flamingo(cfg, dtype=torch.float32, device="cpu", gate=0.3, seed=1)
    sd = get_checkpoint(m, drop_frozen_aliases=True)
    keys = set(sd)
    assert "perceiver.latents" in keys and "perceiver.layers.0.0.to_kv.weight" in keys
    assert "lang_encoder.gated_cross_attn_layers.0.attn.to_q.weight" in keys
    assert "lang_encoder.gated_cross_attn_layers.1.ff_gate" in keys
    assert "lang_encoder.gpt_neox.embed_in.weight" in keys                   # trainable input embeddings
    assert not any(k.startswith("vision_encoder.") for k in keys)             # frozen tower dropped
    assert "lang_encoder.embed_out.weight" not in keys                        # frozen output head dropped
    assert not any("decoder_layer.attention" in k or k.startswith("lang_encoder.old_decoder_blocks") for k in keys)
    # the reference's own behaviour keeps the aliased frozen LM blocks
    assert any(k.startswith("lang_encoder.old_decoder_blocks") for k in get_checkpoint(m))
    m2 = build_flamingo(cfg, dtype=torch.float32, device="cpu", gate=None, seed=2)
    missing, unexpected = m2.load_state_dict(sd, strict=False)
    assert not unexpected
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), m2.named_parameters()):
        if p1.requires_grad:
            assert torch.equal(p1, p2), n1


def test_additive_mask_cache_is_keyed_on_identity_dtype_and_version():
    """`fused_neox_layer` converts HF's boolean attention mask once per forward; the cache must not
    serve a stale conversion after an in-place edit or for another dtype."""
    from unimp_b200.flamingo_lm import _additive_mask

    m = torch.tensor([[True, False, True]])
    a = _additive_mask(m, torch.float32)
    assert _additive_mask(m, torch.float32) is a and a.tolist() == [[0.0, float("-inf"), 0.0]]
    assert _additive_mask(m, torch.bfloat16).dtype == torch.bfloat16
    m[0, 1] = True
    assert _additive_mask(m, torch.float32).tolist() == [[0.0, 0.0, 0.0]]
    assert _additive_mask(m.clone(), torch.float32) is not _additive_mask(m, torch.float32)
    assert _additive_mask(None, torch.float32) is None and _additive_mask(a, torch.float32) is a   # non-bool: as is


def test_resize_token_embeddings_on_the_product_model():
    """reference `UniMP/mmrec.py:595`: `lang_encoder.resize_token_embeddings(len(tokenizer))` after
    the vocabulary grew.  The product's padded output head must survive it: old rows kept, both
    embeddings at the new size, still an nn.Linear, re-wrapped for the new (odd) width, frozen /
    trainable flags kept, state-dict shape exactly (V, D)."""
    import copy
    from unimp_b200.factory import build_flamingo
    from unimp_b200.flamingo_lm import PaddedOutputHead

    cfg = copy.copy(tiny_config())
    cfg.vocab = 515                                    # odd, like the reference's 74 053
    m = build_flamingo(cfg, dtype=torch.float32, device="cpu", seed=3)
    lm = m.lang_encoder
    assert isinstance(lm.embed_out, PaddedOutputHead) and isinstance(lm.embed_out, torch.nn.Linear)
    w_out, w_in = lm.embed_out.weight.detach().clone(), lm.get_input_embeddings().weight.detach().clone()
    lm.resize_token_embeddings(523)
    head = lm.get_output_embeddings()
    assert isinstance(head, PaddedOutputHead) and head.out_features == 523
    assert tuple(head.weight.shape) == (523, cfg.lm_hidden)
    assert tuple(lm.get_input_embeddings().weight.shape) == (523, cfg.lm_hidden)
    assert torch.equal(head.weight[:515], w_out) and torch.equal(lm.get_input_embeddings().weight[:515], w_in)
    assert not head.weight.requires_grad and lm.get_input_embeddings().weight.requires_grad
    sd = m.state_dict()
    assert tuple(sd["lang_encoder.embed_out.weight"].shape) == (523, cfg.lm_hidden)
    lm.resize_token_embeddings(520)                    # multiple of 8: a plain Linear is enough
    assert type(lm.get_output_embeddings()) is torch.nn.Linear
    assert torch.equal(lm.get_output_embeddings().weight[:515], w_out)


def test_adamw_hyper_staging_ring_never_overwrites_an_in_flight_buffer():
    """one reused pinned buffer + non_blocking copy let step k run with step k+n's lr
    and bias corrections.  Every prepare_step must stage through a different buffer than the
    previous HYPER_RING-1 calls."""
    net = torch.nn.Linear(8, 8)
    opt = FlatAdamW([{"params": list(net.named_parameters()), "weight_decay": 0.0}], lr=1e-3)
    seen = []
    for s in range(1, 2 * opt.HYPER_RING + 1):
        opt.prepare_step(lr_scale=s)
        slot = opt.step_count % opt.HYPER_RING
        seen.append(slot)
        assert opt._hyper_ring[slot][0].item() == pytest.approx(1e-3 * s)
        assert opt.hyper[0].item() == pytest.approx(1e-3 * s)
        assert opt.hyper[1].item() == pytest.approx(1 - 0.9 ** s, rel=1e-6)
    for i in range(opt.HYPER_RING, len(seen)):
        assert len(set(seen[i - opt.HYPER_RING + 1:i + 1])) == opt.HYPER_RING   # no reuse inside the window


def test_cached_media_kv_is_rebuilt_for_new_media_and_after_an_optimizer_step():
    """the decode-time to_kv(media) cache was keyed on batch size only."""
    from unimp_b200 import helpers, ops

    attn = helpers.MaskedCrossAttention(dim=32, dim_visual=16)
    calls = []
    attn.project_media = lambda media: calls.append(media) or torch.zeros(1)
    a, b = torch.randn(2, 2, 64, 16), torch.randn(2, 2, 64, 16)
    attn.cached_media_kv(a); attn.cached_media_kv(a)
    assert len(calls) == 1
    attn.cached_media_kv(b)                            # same batch size, new images
    assert len(calls) == 2
    ops.bump_weights_epoch()                           # an optimizer step (raw-pointer update)
    attn.cached_media_kv(b)
    assert len(calls) == 3
    b.add_(1.0)                                        # media edited in place
    attn.cached_media_kv(b)
    assert len(calls) == 4


def test_direct_grad_parameter_fails_loudly_when_autograd_delivers_its_gradient():
    """a direct-accumulation parameter reached through plain autograd used to lose its
    gradient silently."""
    class Blk(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.to_q = torch.nn.Linear(4, 4, bias=False)
    net = Blk()
    opt = FlatAdamW([{"params": list(net.named_parameters()), "weight_decay": 0.0}], lr=1e-3)
    assert not opt._guards                             # CPU tensors are never "direct"
    net.to_q.weight._unimp_direct = True               # emulate the CUDA registration
    opt.groups[0]["n_direct"] = 1
    opt._install_direct_guards()
    with pytest.raises(RuntimeError, match="direct gradient accumulation"):
        net.to_q(torch.randn(2, 4)).sum().backward()
    # the sanctioned route (backward returns None for the weight) must NOT trip the guard
    from unimp_b200 import ops
    net.to_q.weight._unimp_fresh = True
    ops.linear_acc(torch.randn(2, 4, requires_grad=True), net.to_q.weight).sum().backward()
    assert net.to_q.weight._unimp_fresh is False
