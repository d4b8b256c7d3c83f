"""Generates tests/golden/ref_*.pt by RUNNING THE REFERENCE'S OWN CODE (in this container only).

    python tests/golden/make_reference_golden.py          # needs /root/reference

What is executed is the unmodified reference, imported from where it lies:

* `UniMP/mmrec.py::train_one_epoch` (`:65-302`) — the real training-loop body: batch unpack
  (`:135-141`), the answer-span label state machine (`:143-168`), the model call, the task-weighted
  focal loss (`:190-213`) and `accelerator.backward`.  The module's missing third-party imports
  (`open_flamingo`, `accelerate`, `wandb`, `webdataset`, ...) are replaced by empty
  stub modules — none of them is touched by `train_one_epoch` — and the loop is driven with a
  recording model (returns prescribed logits, records the labels it was handed), a recording
  accelerator (`backward` keeps the loss and calls `loss.backward()`), and no-op optimizer /
  scheduler.  What comes out — labels, loss, dloss/dlogits — IS the reference's arithmetic for the
  in-tree half of the path.
* `UniMP/pipeline/mm_utils/collate_rec.py::collate_fn` (`:38-74`) — the batch layout.
* `UniMP/pipeline/train/train_utils.py::get_checkpoint` (`:258-265`).
* `get_grouped_params` / `apply_decay` (`UniMP/mmrec.py:609-631`) — nested inside `main()`, so its
  FunctionDef node is lifted out of the parsed source with `ast` and compiled as is.

* `UniMP/xformers_model/clip.py::CLIPVisionModel` (`:488-543`; attention `:88-141`) — the in-tree
  ViT tower.  Its only missing dependency is `xformers.ops.memory_efficient_attention`, for which a
  dense statement of that op's documented contract is injected (inputs `(B, M, H, K)`,
  `softmax(q kT * scale) v`); embeddings, pre-LN, blocks, quick-GELU MLP and the
  `last_hidden_state` convention are the reference's code.

* `UniMP/pipeline/mm_utils/rec_dataset.py::RecDataset.process_train_{rec,rate_exp,img_gen}_pair`
  (`:372-456`, `:1100-1156`, `:719-777`) with `extract_meta` / `extract_meta_gen` (`:301-337`) — the
  prompt templates — called unbound on a stand-in `self` (fake item metadata, tiny JPEGs, a
  word-level tokenizer that splits the added tokens out the way HF does), then the reference's
  `collate_fn` and the label masking of `train_one_epoch`: the token GRAMMAR of real batches, which
  `unimp_b200/synth.py` must reproduce.

* the vocabulary growth of `main()` (`UniMP/mmrec.py:538-581`: `<answer>`, `rate_*`, `s_*`, `item_*`,
  `img_*,` appended to the tokenizer in that order) — the statements are lifted out of `main()` by
  line range with `ast` and executed against a recording tokenizer.

No reference source is copied into the repo; only the small input/output vectors are committed.
The third-party half (open_flamingo v2.0.1) is absent from /root/reference and stays unpinned.
"""
import argparse
import ast
import contextlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/UniMP"

# modules the reference imports at module top that this image does not have (or that drag in
# datasets / webdataset); train_one_epoch / collate_fn / get_checkpoint use none of them
STUB_PREFIXES = ("wandb", "open_flamingo", "accelerate", "webdataset", "braceexpand", "deepspeed",
                 "pipeline.train.data", "pipeline.eval")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(STUB_PREFIXES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        module.__path__ = []


def import_reference():
    import transformers  # noqa: real package, imported BEFORE the stub finder so its own
    import transformers.modeling_outputs  # noqa  optional-dependency probes see the real environment
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF)
    import mmrec  # noqa: the reference's training script, unmodified
    from pipeline.mm_utils.collate_rec import collate_fn
    from pipeline.train.train_utils import get_checkpoint
    return mmrec, collate_fn, get_checkpoint


# ------------------------------------------------------------------------------------------------
class Tok:
    """Callable like the HF tokenizer at `mmrec.py:84-88`: tok(text)["input_ids"][-1]."""

    def __init__(self, ids, pad):
        self.ids, self.pad_token_id = ids, pad

    def __call__(self, text, add_special_tokens=False):
        return {"input_ids": [self.ids[text]]}


class RecordingModel(torch.nn.Module):
    def __init__(self, logits):
        super().__init__()
        self.logits = torch.nn.Parameter(logits.clone())
        self.seen = {}

    def forward(self, vision_x=None, lang_x=None, attention_mask=None, labels=None):
        from transformers.modeling_outputs import CausalLMOutputWithPast
        self.seen = {"vision_x_shape": tuple(vision_x.shape), "labels": labels.clone()}
        return CausalLMOutputWithPast(loss=self.logits.sum() * 0.0, logits=self.logits)


class RecordingAccelerator:
    sync_gradients = True

    def __init__(self):
        self.loss = None

    @contextlib.contextmanager
    def accumulate(self, model):
        yield

    def backward(self, loss):
        self.loss = loss.detach().clone()
        loss.backward()

    def clip_grad_norm_(self, params, max_norm):
        return None


class _Noop:
    param_groups = [{"lr": 0.0}]

    def step(self):
        pass

    def zero_grad(self):
        pass


def run_reference_step(mmrec, batch, logits, tok_ids, pad, *, gamma, use_reweight):
    args = argparse.Namespace(num_epochs=1, precision="fp32", task="rec", gamma=gamma, rank=0,
                              use_reweight=use_reweight, mask_lm_head=False,
                              gradient_accumulation_steps=1, batch_size=batch["input_ids"].shape[0],
                              world_size=1, report_to_wandb=False, logging_steps=10 ** 9)
    model, acc = RecordingModel(logits), RecordingAccelerator()
    loader = [{"net_input": {k: v.clone() for k, v in batch.items()}}]
    mmrec.train_one_epoch(args, model, 0, loader, Tok(tok_ids, pad), _Noop(), _Noop(), 0, acc, None)
    return {"labels": model.seen["labels"], "vision_x_shape": model.seen["vision_x_shape"],
            "loss": acc.loss, "dlogits": model.logits.grad.clone()}


def lifted_get_grouped_params(weight_decay):
    """`get_grouped_params` is a closure inside `main()` (`mmrec.py:609-631`): lift its FunctionDef
    out of the parsed module and compile it unchanged; `args` is the only free variable."""
    src = open(os.path.join(REF, "mmrec.py")).read()
    tree = ast.parse(src)
    node = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_grouped_params")
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {"args": argparse.Namespace(weight_decay=weight_decay)}
    exec(compile(mod, "mmrec.py::get_grouped_params", "exec"), ns)
    return ns["get_grouped_params"], node.lineno, node.end_lineno


def reference_vit_golden(cfg, g):
    """Runs the reference's in-tree CLIP vision tower on the C1 tiny shape."""
    xf, xo = types.ModuleType("xformers"), types.ModuleType("xformers.ops")

    def memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
        assert attn_bias is None and p == 0.0
        scale = q.shape[-1] ** -0.5 if scale is None else scale
        a = (torch.einsum("bmhk,bnhk->bhmn", q.double(), k.double()) * scale).softmax(-1)
        return torch.einsum("bhmn,bnhk->bmhk", a, v.double()).to(q.dtype)

    xo.memory_efficient_attention = memory_efficient_attention
    xo.LowerTriangularMask = type("LowerTriangularMask", (), {})
    xf.ops = xo
    sys.modules["xformers"], sys.modules["xformers.ops"] = xf, xo
    from transformers import CLIPVisionConfig
    from xformers_model import clip as ref_clip  # the reference's file, unmodified

    vc = CLIPVisionConfig(hidden_size=cfg.vis_width, num_hidden_layers=cfg.vis_layers,
                          num_attention_heads=cfg.vis_heads, intermediate_size=cfg.vis_mlp,
                          image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act="quick_gelu")
    model = ref_clip.CLIPVisionModel(vc).eval()
    # weights on a 1/256 grid (exact in fp32 and bf16), stored as int8; LN weights around 1
    sd_i8 = {}
    for k, v in model.state_dict().items():
        if not v.is_floating_point():
            continue
        q = torch.randint(-64, 64, v.shape, generator=g, dtype=torch.int8)
        sd_i8[k] = q
        base = 1.0 if (k.endswith("norm.weight") or "layer_norm" in k and k.endswith(".weight")
                       or k.endswith("layrnorm.weight")) else 0.0
        v.copy_(base + q.float() / 256.0)
    pixels_i8 = torch.randint(-96, 96, (3, 3, cfg.image_size, cfg.image_size), generator=g, dtype=torch.int8)
    with torch.no_grad():
        out = model(pixel_values=pixels_i8.float() / 32.0).last_hidden_state
    return {"state_i8": sd_i8, "state_scale": 1.0 / 256.0, "pixels_i8": pixels_i8, "pixel_scale": 1.0 / 32.0,
            "last_hidden_state": out,
            "source": "UniMP/xformers_model/clip.py::CLIPVisionModel executed unmodified; "
                      "xformers.ops.memory_efficient_attention replaced by its dense definition"}


def lifted_vocab_growth(subset="all", use_semantic=False):
    """Executes the `tokenizer.add_special_tokens / add_tokens` block of `main()` unmodified and
    returns the tokens in the order the reference appends them."""
    src = open(os.path.join(REF, "mmrec.py")).read()
    tree = ast.parse(src)
    main_fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    seg = lambda n: ast.get_source_segment(src, n) or ""
    first = next(i for i, n in enumerate(main_fn.body) if "tokenizer.add_special_tokens" in seg(n))
    last = next(i for i, n in enumerate(main_fn.body) if "resize_token_embeddings" in seg(n))
    block = main_fn.body[first:last]
    added = []

    class Rec:
        def add_special_tokens(self, d):
            added.extend(d["additional_special_tokens"])

        def add_tokens(self, toks):
            added.extend(toks)

    ns = {"tokenizer": Rec(), "args": argparse.Namespace(subset=subset, use_semantic=use_semantic)}
    exec(compile(ast.Module(body=block, type_ignores=[]), "mmrec.py::main[vocab growth]", "exec"), ns)
    return added, block[0].lineno, block[-1].end_lineno


class WordTokenizer:
    """Stand-in for the HF tokenizer the reference builds (`mmrec.py:538-595`): added tokens
    (`<image>`, `<answer>`, `<|endofchunk|>`, `item_N`, `img_N`, `rate_N`) are matched first, the rest
    is split on whitespace / punctuation and hashed into the plain-text id range."""

    def __init__(self, tk):
        import re
        self.tk = tk
        self.rx = re.compile(r"<image>|<answer>|<\|endofchunk\|>|item_\d+|img_\d+|rate_\d+|[A-Za-z0-9]+|[^\sA-Za-z0-9]")

    def ids(self, text):
        tk, out = self.tk, []
        for w in self.rx.findall(text):
            if w == "<image>":
                out.append(tk.media)
            elif w == "<answer>":
                out.append(tk.answer)
            elif w == "<|endofchunk|>":
                out.append(tk.endofchunk)
            elif w.startswith("item_"):
                out.append(tk.first_item + int(w[5:]) % tk.n_items)
            elif w.startswith("img_"):
                out.append(tk.first_img + int(w[4:]) % tk.n_img)
            elif w.startswith("rate_"):
                out.append(tk.first_item + tk.n_items - 1 - int(w[5:]) % 6)   # some added-token id
            else:
                out.append(1 + sum(ord(c) * (i + 7) for i, c in enumerate(w)) % (tk.n_plain - 1))
        return out

    def __call__(self, text, return_tensors="pt", add_special_tokens=False, truncation=False):
        ids = torch.tensor([self.ids(text)], dtype=torch.int64)
        return {"input_ids": ids, "attention_mask": torch.ones_like(ids)}


def reference_dataset_batches(mmrec, collate_fn, cfg, tok_ids):
    """Real batch structure: RecDataset sample builders -> collate_fn -> train_one_epoch labels."""
    import tempfile
    import types as _t

    import numpy as np
    from PIL import Image
    from pipeline.mm_utils.rec_dataset import RecDataset
    from torchvision import transforms

    tk = cfg.tokens
    np.random.seed(7)
    tmp = tempfile.mkdtemp(prefix="unimp_golden_")
    items = list(range(1, 13))
    for it in items:
        Image.fromarray((np.random.rand(10, 10, 3) * 255).astype("uint8")).save(os.path.join(tmp, f"{it}.jpg"))
    meta = {str(it): {"category": f"Cat{it % 3} sub", "brand": "" if it % 4 == 0 else f"Brand{it}",
                      "title": f"Title of item {it} with words", "price": "" if it % 5 == 0 else f"{it}.99",
                      "retrieval": [items[(it + 3) % len(items)]], "keywords": f"kw{it} red shoes"}
            for it in items}
    ds = _t.SimpleNamespace(
        seqs=[[(it, f"great item {it} really", 1 + it % 5) for it in items[s:s + 6]] for s in range(0, 6)],
        history_len=2, img_folder=tmp, subset="all", use_semantic=False, meta_data=meta,
        img_id2semantic={str(it): [it, it + 1, it + 2, it + 3] for it in items},
        patch_resize_transform=transforms.Compose([transforms.Resize((8, 8)), transforms.ToTensor()]),
        tokenizer=WordTokenizer(tk),
        bos_item=torch.LongTensor([tk.bos]), eos_item=torch.LongTensor([tk.eos]),
        bos_mask=torch.LongTensor([1]), eos_mask=torch.LongTensor([1]))
    ds.extract_meta = lambda i: RecDataset.extract_meta(ds, i)
    ds.extract_meta_gen = lambda i: RecDataset.extract_meta_gen(ds, i)
    out = {}
    for task, fn in (("rec", RecDataset.process_train_rec_pair), ("rate_exp", RecDataset.process_train_rate_exp_pair),
                     ("img_gen", RecDataset.process_train_img_gen_pair)):
        samples = [fn(ds, i) for i in range(3)]
        batch = collate_fn(samples, pad_idx=tk.pad, eos_idx=tk.eos)["net_input"]
        B, T = batch["input_ids"].shape
        step = run_reference_step(mmrec, batch, torch.zeros(B, T, cfg.vocab), tok_ids, tk.pad, gamma=2.0,
                                  use_reweight=True)
        out[task] = {"input_ids": batch["input_ids"], "attention_masks": batch["attention_masks"],
                     "weights": batch["weights"], "n_images": batch["patch_images"].shape[1],
                     "labels": step["labels"]}
        print("dataset", task, tuple(batch["input_ids"].shape), "images", batch["patch_images"].shape[1],
              "weights", batch["weights"].tolist(), "valid labels", int((step["labels"] != -100).sum()))
    return out


def main():
    torch.set_num_threads(1)
    sys.path.insert(0, ROOT)
    from unimp_b200 import tiny_config
    from unimp_b200.config import WORKLOADS, Workload
    from unimp_b200.synth import make_batch

    mmrec, collate_fn, get_checkpoint = import_reference()
    cfg = tiny_config()
    tk = cfg.tokens
    tok_ids = {"<image>": tk.media, "<|endofchunk|>": tk.endofchunk, "<answer>": tk.answer}
    V = cfg.vocab
    g = torch.Generator().manual_seed(4321)

    # ---- train-step cases: labels + focal loss + its gradient ---------------------------------
    cases = []
    b1 = make_batch(cfg, WORKLOADS["C1-tiny"], seed=1234, ragged=True)
    b2 = make_batch(cfg, Workload("golden-b3", B=3, Ti=2, T=48, row_weights=(2.0, 1.0)), seed=77)
    # hand-written adversarial rows for the state machine (mmrec.py:146-156): <answer> twice in a
    # row, <|endofchunk|> outside an answer, pad inside an answer, <image> inside an answer,
    # an answer that never closes, <answer> at position 0
    A, E, M, P = tk.answer, tk.endofchunk, tk.media, tk.pad
    adv = torch.tensor([
        [1, 5, A, A, 7, E, E, 9, A, 11, P, 12, E, M, 13, A],
        [A, 3, 4, E, M, 5, A, 6, M, 7, 8, 9, 10, 11, 12, 13],
        [1, E, 2, M, 3, 4, 5, 6, 7, 8, 9, A, 10, 2, P, P],
    ], dtype=torch.int64)
    b3 = {"input_ids": adv, "attention_masks": (adv != P).long(),
          "patch_images": torch.zeros(3, 1, 3, 2, 2), "weights": torch.tensor([1.0, 2.0, 0.5])}
    for name, batch, gamma, use in [("c1_ragged_g2", b1, 2.0, True), ("b3_g2", b2, 2.0, True),
                                    ("b3_g0.5", b2, 0.5, True), ("b3_plain", b2, 2.0, False),
                                    ("adversarial_g2", b3, 2.0, True)]:
        B, T = batch["input_ids"].shape
        # logits on a 1/8 grid in [-6, 6): exact in fp32 AND bf16, stored as int8 (small fixture)
        logits_i8 = torch.randint(-48, 48, (B, T, V), generator=g, dtype=torch.int8)
        logits = logits_i8.float() * 0.125
        out = run_reference_step(mmrec, batch, logits, tok_ids, tk.pad, gamma=gamma, use_reweight=use)
        assert out["vision_x_shape"][2] == 1            # unsqueeze(2): (B, Ti, 1, C, H, W)
        cases.append({"name": name, "input_ids": batch["input_ids"], "weights": batch["weights"],
                      "logits_i8": logits_i8, "logits_scale": 0.125, "gamma": gamma, "use_reweight": use,
                      "labels": out["labels"], "loss": out["loss"],
                      # dense gradient is zero off the valid rows: keep only those
                      "dlogits_rows": out["dlogits"].flatten(0, 1).abs().sum(-1).nonzero().flatten(),
                      "dlogits_vals": out["dlogits"].flatten(0, 1)[
                          out["dlogits"].flatten(0, 1).abs().sum(-1).nonzero().flatten()],
                      "vision_x_shape": out["vision_x_shape"]})
        print(name, "loss", float(out["loss"]), "n_valid", int((out["labels"][:, 1:] != -100).sum()))
    torch.save({"tokens": {"answer": A, "endofchunk": E, "media": M, "pad": P}, "cases": cases,
                "source": "UniMP/mmrec.py::train_one_epoch executed unmodified"},
               os.path.join(HERE, "ref_train_step.pt"))

    # ---- the in-tree ViT tower ---------------------------------------------------------------
    vit = reference_vit_golden(cfg, g)
    torch.save(vit, os.path.join(HERE, "ref_vit_tower.pt"))
    print("vit tokens", tuple(vit["last_hidden_state"].shape), "rms", float(vit["last_hidden_state"].pow(2).mean().sqrt()))

    # ---- real batch structure from the dataset code -------------------------------------------
    ds = reference_dataset_batches(mmrec, collate_fn, cfg, tok_ids)
    torch.save({"tokens": {"answer": A, "endofchunk": E, "media": M, "pad": P, "bos": tk.bos, "eos": tk.eos,
                           "first_item": tk.first_item, "n_items": tk.n_items, "first_img": tk.first_img,
                           "n_img": tk.n_img, "n_plain": tk.n_plain},
                "tasks": ds,
                "source": "RecDataset.process_train_{rec,rate_exp,img_gen}_pair + collate_fn + train_one_epoch "
                          "label masking, executed unmodified on stand-in data"},
               os.path.join(HERE, "ref_dataset_batches.pt"))

    # ---- collate_fn ------------------------------------------------------------------------------
    lens = [9, 14, 5]
    samples = [{"net_input": {"input_ids": torch.randint(1, 90, (n,), generator=g),
                              "attention_masks": torch.ones(n, dtype=torch.int64),
                              "weights": w,
                              "patch_images": torch.randn(2, 3, 4, 4, generator=g)}}
               for n, w in zip(lens, [2.0, 1.0, 1.0])]
    ref_batch = collate_fn(samples, pad_idx=tk.pad, eos_idx=tk.eos)
    torch.save({"samples": samples, "pad_idx": tk.pad, "eos_idx": tk.eos, "batch": ref_batch,
                "source": "UniMP/pipeline/mm_utils/collate_rec.py::collate_fn executed unmodified"},
               os.path.join(HERE, "ref_collate.pt"))

    # ---- get_checkpoint + weight-decay groups on the real tiny module tree ------------------------
    names = [
        "perceiver.latents", "perceiver.layers.0.0.norm_media.weight", "perceiver.layers.0.0.to_q.weight",
        "perceiver.layers.0.1.0.bias", "perceiver.norm.weight",
        "lang_encoder.gated_cross_attn_layers.0.attn_gate", "lang_encoder.gated_cross_attn_layers.0.ff_gate",
        "lang_encoder.gated_cross_attn_layers.0.attn.norm.weight",
        "lang_encoder.gated_cross_attn_layers.0.attn.norm.bias",
        "lang_encoder.gated_cross_attn_layers.0.attn.to_q.weight",
        "lang_encoder.gated_cross_attn_layers.0.attn.to_kv.weight",
        "lang_encoder.gated_cross_attn_layers.0.attn.to_out.weight",
        "lang_encoder.gated_cross_attn_layers.0.ff.0.weight", "lang_encoder.gated_cross_attn_layers.0.ff.0.bias",
        "lang_encoder.gated_cross_attn_layers.0.ff.1.weight", "lang_encoder.gated_cross_attn_layers.0.ff.3.weight",
        "lang_encoder.gpt_neox.embed_in.weight", "lang_encoder.embed_out.weight",
        "lang_encoder.gpt_neox.layers.0.gated_cross_attn_layer.attn.to_q.weight",
        "lang_encoder.gpt_neox.layers.0.gated_cross_attn_layer.ff.0.weight",
        "lang_encoder.gpt_neox.layers.0.decoder_layer.attention.dense.weight",
        "vision_encoder.transformer.resblocks.0.ln_1.weight",
    ]

    class _Named(torch.nn.Module):
        def __init__(self, names):
            super().__init__()
            self._n = names
            self._p = [torch.nn.Parameter(torch.zeros(1)) for _ in names]

        def named_parameters(self, *a, **k):
            return iter(zip(self._n, self._p))

    ggp, l0, l1 = lifted_get_grouped_params(0.1)
    m = _Named(names)
    groups = ggp(m)
    wd_ids = {id(p) for p in groups[0]["params"]}
    decay = {n: (id(p) in wd_ids) for n, p in zip(m._n, m._p)}
    assert groups[0]["weight_decay"] == 0.1 and groups[1]["weight_decay"] == 0.0

    class _Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.frozen = torch.nn.Linear(3, 3)
            self.train_me = torch.nn.Linear(3, 2, bias=False)
            self.alias = torch.nn.ModuleList([self.frozen])   # same tensors under a 2nd name
            self.register_buffer("buf", torch.zeros(2))
            for p in self.frozen.parameters():
                p.requires_grad_(False)

    import hashlib
    vocab_added, v0, v1 = lifted_vocab_growth("all", False)
    groups, cur = [], None
    for i, t in enumerate(vocab_added):           # runs of tokens sharing a prefix
        pre = t.rstrip(",0123456789")
        if cur is None or cur[0] != pre:
            cur = [pre, i, 0, t, t]
            groups.append(cur)
        cur[2] += 1
        cur[4] = t
    vocab = {"lines": (v0, v1), "n_added": len(vocab_added), "groups": [tuple(g) for g in groups],
             "sha256": hashlib.sha256("\n".join(vocab_added).encode()).hexdigest()}
    print("vocab growth lines", v0, v1, "added", len(vocab_added), [(g[0], g[2]) for g in groups])

    ck = get_checkpoint(_Toy())
    torch.save({"decay": decay, "vocab_growth": vocab, "grouped_params_lines": (l0, l1), "toy_checkpoint_keys": sorted(ck.keys()),
                "source": "mmrec.py::get_grouped_params (ast-lifted) and train_utils.py::get_checkpoint, unmodified"},
               os.path.join(HERE, "ref_host_rules.pt"))
    print("decay", sum(decay.values()), "of", len(decay), "| toy checkpoint keys", sorted(ck.keys()))


if __name__ == "__main__":
    main()
