"""Decoding rules of unimp_b200/decode.py against transformers' own `generate` (CPU), and the
CUDA-graph decoder against `Flamingo.generate` (GPU).

The reference decodes with HF beam search (`UniMP/pipeline/eval/eval_exp.py:101-114`: num_beams=5,
early_stopping=True, num_return_sequences=K, max_new_tokens=256, eos/pad ids).  `BeamSearch` /
`GreedySearch` restate those rules with device-side cursors; here they drive a plain HF GPT-NeoX
through a cache-free step function and must return exactly the token matrix HF returns.
"""
import pytest
import torch

from unimp_b200.decode import generate_with


def _tiny_lm(vocab, seed, eos_boost=1.2):
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

    torch.manual_seed(seed)
    cfg = GPTNeoXConfig(hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                        vocab_size=vocab, max_position_embeddings=128, tie_word_embeddings=False,
                        hidden_dropout=0.0, attention_dropout=0.0)
    m = GPTNeoXForCausalLM(cfg).eval()
    with torch.no_grad():   # sharpen the next-token distributions so that beams differ ...
        m.embed_out.weight.mul_(8.0)
        # ... and give EOS (id 1) a steady boost through the final LayerNorm's bias direction, so that
        # hypotheses do finish: the finished-pool / early-stopping rules are what is under test
        b = torch.randn(32)
        m.gpt_neox.final_layer_norm.bias.copy_(0.3 * b)
        m.embed_out.weight[1] += eos_boost * b / (0.3 * b.pow(2).sum())
    return m


@pytest.mark.parametrize("nb,nrs,early,lp", [(1, 1, False, 1.0), (3, 1, True, 1.0), (5, 3, True, 1.0),
                                             (4, 4, False, 1.0), (4, 2, "never", 1.0), (3, 2, True, 0.5),
                                             (5, 5, False, 2.0)])
@pytest.mark.parametrize("vocab,seed", [(13, 0), (40, 1), (97, 2)])
def test_search_rules_equal_transformers_generate(nb, nrs, early, lp, vocab, seed):
    torch.set_num_threads(1)
    m = _tiny_lm(vocab, seed)
    g = torch.Generator().manual_seed(100 + seed)
    prompt = torch.randint(3, vocab, (2, 6), generator=g)
    eos, pad, new = 1, 2, 12
    kw = dict(max_new_tokens=new, eos_token_id=eos, pad_token_id=pad, do_sample=False)
    if nb > 1:
        kw.update(num_beams=nb, num_return_sequences=nrs, early_stopping=early, length_penalty=lp)
    with torch.no_grad():
        want = m.generate(prompt, attention_mask=torch.ones_like(prompt), **kw)

    def step_fn(rows):
        return m(input_ids=rows).logits[:, -1, :]

    got = generate_with(step_fn, prompt, num_beams=nb, max_new_tokens=new, eos_token_id=eos, pad_token_id=pad,
                        num_return_sequences=nrs, early_stopping=early, length_penalty=lp)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert torch.equal(got, want)


@pytest.mark.parametrize("B,T0,new,nb", [(1, 3, 1, 2), (1, 9, 20, 5), (3, 4, 6, 3), (3, 5, 2, 1), (4, 2, 15, 2)])
def test_search_rules_other_batch_sizes_prompt_lengths_and_budgets(B, T0, new, nb):
    """Batch of one, three-row batches (rows finish at different steps), a single new token, long budgets."""
    torch.set_num_threads(1)
    m = _tiny_lm(29, 7)
    g = torch.Generator().manual_seed(B * 100 + T0)
    prompt = torch.randint(3, 29, (B, T0), generator=g)
    kw = dict(max_new_tokens=new, eos_token_id=1, pad_token_id=2, do_sample=False)
    if nb > 1:
        kw.update(num_beams=nb, num_return_sequences=nb, early_stopping=True)
    with torch.no_grad():
        want = m.generate(prompt, attention_mask=torch.ones_like(prompt), **kw)
    got = generate_with(lambda rows: m(input_ids=rows).logits[:, -1, :], prompt, num_beams=nb, max_new_tokens=new,
                        eos_token_id=1, pad_token_id=2, num_return_sequences=nb, early_stopping=True)
    assert got.shape == want.shape and torch.equal(got, want)


def test_eos_is_actually_exercised():
    """Guard for the test above: with these tiny models some hypotheses do end in EOS and some runs
    stop before max_new_tokens (otherwise the finished-pool rules would go untested)."""
    ended, short = 0, 0
    for vocab, seed in [(13, 0), (40, 1), (97, 2)]:
        m = _tiny_lm(vocab, seed)
        g = torch.Generator().manual_seed(100 + seed)
        prompt = torch.randint(3, vocab, (2, 6), generator=g)
        out = generate_with(lambda rows: m(input_ids=rows).logits[:, -1, :], prompt, num_beams=5,
                            max_new_tokens=12, eos_token_id=1, pad_token_id=2, num_return_sequences=5,
                            early_stopping=True)
        ended += int((out[:, 6:] == 1).any(-1).sum())
        short += int(out.shape[1] < 18)
    assert ended > 0 and short > 0


# ------------------------------------------------------------------------------- GPU: the graph decoder

def _tiny_flamingo(dtype):
    from unimp_b200 import tiny_config
    from unimp_b200.config import Workload
    from unimp_b200.factory import build_flamingo
    from unimp_b200.synth import make_batch

    cfg = tiny_config()
    model = build_flamingo(cfg, dtype=dtype, device="cuda", gate=0.5).eval()
    b = make_batch(cfg, Workload("dec", B=2, Ti=2, T=40), seed=3)
    L = int(b["attention_masks"].sum(-1).min()) - 2
    ids = b["input_ids"][:, :L].cuda()
    vis = b["patch_images"].unsqueeze(2).to("cuda", dtype)
    return cfg, model, ids, vis


@pytest.mark.gpu
@pytest.mark.parametrize("nb,nrs,early,new", [(1, 1, False, 9), (3, 2, True, 9), (5, 5, True, 12), (4, 1, False, 7),
                                              (2, 2, True, 1), (3, 1, True, 2)])
def test_graphed_decoder_equals_flamingo_generate(nb, nrs, early, new):
    """CUDA-graph decode (static K/V caches, cached to_kv(media), unimp_xattn_decode, in-graph beam
    search) returns the token matrix of `Flamingo.generate` (HF generate over the same modules)."""
    from unimp_b200.decode import GraphedDecoder

    cfg, model, ids, vis = _tiny_flamingo(torch.float32)
    kw = dict(max_new_tokens=new, eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad)
    if nb > 1:
        kw.update(num_beams=nb, num_return_sequences=nrs, early_stopping=early)
    want = model.generate(vision_x=vis, lang_x=ids, attention_mask=torch.ones_like(ids), do_sample=False, **kw)
    dec = GraphedDecoder(model, sync_every=4)
    got = dec.generate(vis, ids, torch.ones_like(ids), num_beams=nb, max_new_tokens=new,
                       eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad,
                       num_return_sequences=nrs, early_stopping=early)
    assert got.shape == want.shape and torch.equal(got, want)
    # the decoder leaves the model as it found it: a training-mode forward still works
    assert all(l.vis_x is None for l in model.lang_encoder._get_decoder_layers())


@pytest.mark.gpu
def test_graphed_decoder_beam_indirection_equals_reordered_caches():
    """The default decoder never copies K/V caches between beams (an indirection table read by
    `unimp_lm_decode_attn`); the legacy path (HF `reorder_cache` semantics: gather every layer's K/V,
    SDPA over them) must return the same token matrix."""
    from unimp_b200.decode import GraphedDecoder

    cfg, model, ids, vis = _tiny_flamingo(torch.float32)
    kw = dict(num_beams=5, max_new_tokens=14, eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad,
              num_return_sequences=5, early_stopping=False)
    new = GraphedDecoder(model)
    assert new.use_indirection
    old = GraphedDecoder(model)
    old.use_indirection = False
    assert torch.equal(new.generate(vis, ids, None, **kw), old.generate(vis, ids, None, **kw))


@pytest.mark.gpu
def test_graphed_decoder_ragged_prompts_and_bf16():
    from unimp_b200.decode import GraphedDecoder

    cfg, model, ids, vis = _tiny_flamingo(torch.float32)
    ids2, m2 = ids.clone(), torch.ones_like(ids)
    ids2[1, -5:], m2[1, -5:] = cfg.tokens.pad, 0          # right padding, as the reference's collate pads
    kw = dict(max_new_tokens=8, eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad, num_beams=3,
              num_return_sequences=2, early_stopping=True)
    want = model.generate(vision_x=vis, lang_x=ids2, attention_mask=m2, do_sample=False, **kw)
    got = GraphedDecoder(model).generate(vis, ids2, m2, **kw)
    assert torch.equal(got, want)
    # bf16 (the deployment dtype): same shapes, prompt preserved, tokens in range (near-ties may
    # legitimately resolve differently from the eager path at bf16 resolution)
    cfg, model, ids, vis = _tiny_flamingo(torch.bfloat16)
    out = GraphedDecoder(model).generate(vis, ids, None, num_beams=5, max_new_tokens=16, num_return_sequences=5,
                                         eos_token_id=cfg.tokens.endofchunk, pad_token_id=cfg.tokens.pad,
                                         early_stopping=True)
    assert out.shape[0] == 10 and torch.equal(out[:, :ids.shape[1]], ids.repeat_interleave(5, 0))
    assert int(out.min()) >= 0 and int(out.max()) < cfg.vocab


def test_graphed_decoder_host_logic_on_cpu_with_test_doubles():
    """Prefill, static K/V caches, key masks, rotary positions, beam re-ordering and step counting
    of `GraphedDecoder`, run in a throw-away subprocess where tests/standins.py replaces the CUDA ops
    by dense PyTorch doubles: token matrices equal `Flamingo.generate` on the same doubles."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "standins.py"), "decode"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DECODE CHECK ALL EQUAL" in r.stdout and "DIFFERENT" not in r.stdout


def test_graphed_decoder_rejects_what_it_does_not_implement():
    """Sampling / n-gram blocking / other HF generate options must fail loudly, not be ignored."""
    from unimp_b200.decode import GraphedDecoder

    dec = GraphedDecoder.__new__(GraphedDecoder)       # argument validation needs no model
    x = torch.zeros(1, 4, dtype=torch.int64)
    for kw in (dict(do_sample=True), dict(no_repeat_ngram_size=3), dict(temperature=0.7), dict(top_k=5)):
        with pytest.raises(NotImplementedError):
            dec.generate(None, x, None, **kw)
    with pytest.raises(ValueError):
        dec.generate(None, x, None, num_beams=2, num_return_sequences=3)
