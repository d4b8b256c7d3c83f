"""Autoregressive decoding with cached vision latents, captured in a CUDA graph.

What the reference does for explanation generation (`UniMP/pipeline/eval/eval_exp.py:101-114`) is
`model.generate(vision_x=, lang_x=, attention_mask=, num_beams=5, num_return_sequences=K,
early_stopping=True, max_new_tokens=256, eos_token_id=, pad_token_id=)`, i.e. upstream
`Flamingo.generate` -> HF `GenerationMixin.generate` (beam search, `DynamicCache`).  That path is
still available as `Flamingo.generate`; on a B200 it is bound by ~20 ms of Python per token, not
by the GPU.  `GraphedDecoder` is the same computation arranged for the hardware:

* prefill once through the normal module path (vision tower, Perceiver, masked cross-attention
  over the prompt), keeping the LM's K/V;
* every later token is ONE replay of a captured CUDA graph: static K/V caches
  (B*beams, H, T_max, dh) written in place at a device-side cursor, `to_kv(media)` projected once
  per block and cached, single-token masked cross-attention by `unimp_xattn_decode`, GPT-NeoX
  layers through `fused_neox_layer` — rotary + cache write + attention in `unimp_lm_decode_attn`,
  every projection streamed once by `unimp_linear_small_m` — and the beam-search bookkeeping
  (log-softmax, top-k) as static-shape tensor ops inside the same graph.  Beam re-ordering moves the
  rows of an INDIRECTION table (cache row per beam and position), never the caches.  The host only
  launches the graph and reads one "unfinished" flag every `sync_every` tokens.

`BeamSearch` / `GreedySearch` restate the decoding rules of HF `GenerationMixin._beam_search` /
`_sample` (transformers v5: 2*num_beams candidates per step, finished-hypothesis pool with
length penalty, `early_stopping` in {True, False, "never"}, pad after EOS) with device-side
cursors instead of Python integers so that a step is capturable.  They are pure tensor code and are
tested on CPU against `transformers`' own `generate` (tests/test_decode.py).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import ops

NEG = -1.0e9


def _gather(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """t (B, K, ...), idx (B, K') -> t[b, idx[b, k'], ...]"""
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


class GreedySearch:
    """argmax decoding with HF `_sample` semantics: a finished row keeps emitting `pad`."""

    def __init__(self, prompt_ids: torch.Tensor, *, max_new_tokens: int, eos_token_id, pad_token_id):
        B, T0 = prompt_ids.shape
        dev = prompt_ids.device
        self.B, self.T0, self.Lmax = B, T0, T0 + max_new_tokens
        self.eos = None if eos_token_id is None else int(eos_token_id)
        fill = pad_token_id if pad_token_id is not None else (self.eos if self.eos is not None else 0)
        self.pad = int(fill)
        self.seq = torch.full((B, self.Lmax), self.pad, dtype=torch.int64, device=dev)
        self.seq[:, :T0] = prompt_ids
        self.cur = torch.full((1,), T0, dtype=torch.int64, device=dev)   # next position to fill
        self.alive = torch.ones(B, dtype=torch.bool, device=dev)
        self.row_len = torch.zeros(B, dtype=torch.int64, device=dev)     # generated tokens per row
        self.flag = torch.ones((), dtype=torch.bool, device=dev)

    def step(self, logits: torch.Tensor):
        """logits (B, V) -> (next tokens (B,), cache source rows (B,) = identity)."""
        tok = logits.float().argmax(-1)
        tok = torch.where(self.alive, tok, torch.full_like(tok, self.pad))
        self.seq.index_copy_(1, self.cur, tok.unsqueeze(1))
        self.row_len += self.alive.to(torch.int64)
        self.cur += 1
        if self.eos is not None:
            self.alive &= tok != self.eos
        self.alive &= (self.cur < self.Lmax).expand_as(self.alive)
        self.flag.copy_(self.alive.any())
        return tok, None

    def unfinished(self) -> torch.Tensor:
        return self.flag

    def result(self, num_return_sequences: int = 1) -> torch.Tensor:
        L = self.T0 + int(self.row_len.max())
        return self.seq[:, :L].clone()


class BeamSearch:
    """Deterministic beam search, HF `_beam_search` rules (no sampling, no logits processors)."""

    def __init__(self, prompt_ids: torch.Tensor, *, num_beams: int, max_new_tokens: int, eos_token_id,
                 pad_token_id, length_penalty: float = 1.0, early_stopping=False):
        """prompt_ids (B, T0): ONE row per batch item (beams are expanded here)."""
        B, T0 = prompt_ids.shape
        dev = prompt_ids.device
        nb = num_beams
        self.B, self.nb, self.K, self.T0, self.Lmax = B, nb, 2 * nb, T0, T0 + max_new_tokens
        self.eos = None if eos_token_id is None else int(eos_token_id)
        # HF: `pad_token_id or eos_token_id[0] if eos_token_id is not None else -1` (a pad id of 0 is falsy)
        fill = (pad_token_id or self.eos) if self.eos is not None else -1
        self.length_penalty, self.early_stopping = float(length_penalty), early_stopping
        self.running_seq = torch.full((B, nb, self.Lmax), int(fill), dtype=torch.int64, device=dev)
        self.running_seq[:, :, :T0] = prompt_ids[:, None, :]
        self.seq = self.running_seq.clone()
        self.running_scores = torch.zeros((B, nb), dtype=torch.float32, device=dev)
        self.running_scores[:, 1:] = NEG          # only beam 0 is live at the first step
        self.scores = torch.full((B, nb), NEG, dtype=torch.float32, device=dev)
        self.finished = torch.zeros((B, nb), dtype=torch.bool, device=dev)
        self.gen_len = torch.zeros((B, nb), dtype=torch.int64, device=dev)   # of the finished pool
        self.heur_open = torch.ones((B, 1), dtype=torch.bool, device=dev)
        self.top_mask = torch.cat([torch.ones(nb, dtype=torch.bool), torch.zeros(self.K - nb, dtype=torch.bool)]).to(dev)
        self.cur = torch.full((1,), T0, dtype=torch.int64, device=dev)
        self.row_offset = (torch.arange(B, device=dev) * nb)[:, None]
        self.flag = torch.ones((), dtype=torch.bool, device=dev)

    def step(self, logits: torch.Tensor):
        """logits (B*nb, V) of the running beams -> (next tokens (B*nb,), source rows (B*nb,)):
        row i of the next step continues old row `source[i]` (cache re-ordering)."""
        B, nb, K = self.B, self.nb, self.K
        V = logits.shape[-1]
        cur = self.cur
        if logits.is_cuda and ops.BEAM_TOPK and K <= 16 and K <= V and nb <= 64:
            # log_softmax + running score + top-2*nb in two launches (torch: softmax, add, radix top-k)
            top_lp, top_idx = ops.beam_topk(logits.float(), self.running_scores, nb, K)
        else:
            lp = F.log_softmax(logits.float(), dim=-1).view(B, nb, V) + self.running_scores[:, :, None]
            top_lp, top_idx = torch.topk(lp.view(B, nb * V), k=K)
        src_beam = torch.div(top_idx, V, rounding_mode="floor")
        tok = top_idx - src_beam * V
        cand_seq = _gather(self.running_seq, src_beam)
        cand_seq.index_copy_(2, cur, tok.unsqueeze(-1))
        hits = (cur + 1 >= self.Lmax).expand(B, K).clone()
        if self.eos is not None:
            hits |= tok == self.eos
        # the running beams of the next step: best candidates that did not just stop
        run_lp = top_lp + hits.to(torch.float32) * NEG
        nxt = torch.topk(run_lp, k=nb)[1]
        new_running_seq, new_running_scores = _gather(cand_seq, nxt), _gather(run_lp, nxt)
        source = (_gather(src_beam, nxt) + self.row_offset).reshape(-1)
        next_tok = _gather(tok, nxt).reshape(-1)
        # the finished pool: only the top `nb` candidates may enter it
        just = hits & self.top_mask[None, :]
        n_gen = cur + 1 - self.T0                                      # tokens generated so far
        fin_lp = top_lp / (n_gen.to(torch.float32) ** self.length_penalty)
        if self.early_stopping is True:
            fin_lp = fin_lp + self.finished.all(-1, keepdim=True).to(torch.float32) * NEG
        fin_lp = fin_lp + (~self.heur_open).to(torch.float32) * NEG
        fin_lp = fin_lp + (~just).to(torch.float32) * NEG
        m_seq = torch.cat((self.seq, cand_seq), dim=1)
        m_sc = torch.cat((self.scores, fin_lp), dim=1)
        m_fin = torch.cat((self.finished, just), dim=1)
        m_len = torch.cat((self.gen_len, n_gen.expand(B, K)), dim=1)
        sel = torch.topk(m_sc, k=nb)[1]
        # state lives in fixed buffers (a captured step must read what the previous replay wrote)
        self.running_seq.copy_(new_running_seq)
        self.running_scores.copy_(new_running_scores)
        self.seq.copy_(_gather(m_seq, sel))
        self.scores.copy_(_gather(m_sc, sel))
        self.finished.copy_(_gather(m_fin, sel))
        self.gen_len.copy_(_gather(m_len, sel))
        self.cur += 1
        # can a running beam still beat the worst finished one?
        if self.early_stopping == "never" and self.length_penalty > 0.0:
            best_len = torch.full_like(self.cur, self.Lmax - self.T0)
        else:
            best_len = self.cur - self.T0
        best_running = self.running_scores[:, :1] / (best_len.to(torch.float32) ** self.length_penalty)
        worst_fin = torch.where(self.finished, self.scores.min(dim=1, keepdim=True)[0],
                                torch.full_like(self.scores, NEG))
        self.heur_open &= (best_running > worst_fin).any(-1, keepdim=True)
        open_beam = ~self.finished.all() if self.early_stopping is True else torch.ones_like(self.flag)
        self.flag.copy_(self.heur_open.any() & open_beam & ~hits.all())
        return next_tok, source

    def unfinished(self) -> torch.Tensor:
        return self.flag

    def result(self, num_return_sequences: int = 1) -> torch.Tensor:
        n = num_return_sequences
        seq = self.seq[:, :n].reshape(self.B * n, self.Lmax)
        L = self.T0 + int(self.gen_len[:, :n].max())
        return seq[:, :L].clone()


def make_search(prompt_ids, *, num_beams, max_new_tokens, eos_token_id, pad_token_id, length_penalty=1.0,
                early_stopping=False):
    if num_beams == 1:
        return GreedySearch(prompt_ids, max_new_tokens=max_new_tokens, eos_token_id=eos_token_id,
                            pad_token_id=pad_token_id)
    return BeamSearch(prompt_ids, num_beams=num_beams, max_new_tokens=max_new_tokens, eos_token_id=eos_token_id,
                      pad_token_id=pad_token_id, length_penalty=length_penalty, early_stopping=early_stopping)


@torch.no_grad()
def generate_with(step_fn, prompt_ids, **kw):
    """Reference driver (no graph, any device): `step_fn(all_token_rows (B*nb, L)) -> logits
    (B*nb, V)` of the last position.  Used by the CPU tests against `transformers.generate`."""
    nrs = kw.pop("num_return_sequences", 1)
    s = make_search(prompt_ids, **kw)
    nb = getattr(s, "nb", 1)
    rows = prompt_ids.repeat_interleave(nb, dim=0)
    while bool(s.unfinished()):
        tok, src = s.step(step_fn(rows))
        if src is not None:
            rows = rows[src]
        rows = torch.cat([rows, tok[:, None]], dim=1)
    return s.result(nrs)


# ------------------------------------------------------------------------------------------------
# the Flamingo decoder: prefill through the module path, then graph replays
# ------------------------------------------------------------------------------------------------

class GraphedDecoder:
    """`GraphedDecoder(model).generate(...)`: same arguments and result as `Flamingo.generate` for
    deterministic decoding (num_beams >= 1, do_sample=False, no logits processors).  Needs the
    GPT-NeoX language model (the 4B-instruct configuration) on a CUDA device."""

    # additive-mask value for key slots that are not visible (padding, slots past the cursor)
    MASK_FILL = float("-inf")

    def __init__(self, model, *, sync_every: int = 8):
        from transformers.models.gpt_neox.modeling_gpt_neox import GPTNeoXLayer

        self.model = model
        self.lm = model.lang_encoder
        self.layers = list(self.lm._get_decoder_layers())
        if not all(isinstance(l.decoder_layer, GPTNeoXLayer) for l in self.layers):
            raise TypeError("GraphedDecoder drives GPT-NeoX decoder layers")
        self.sync_every = max(1, int(sync_every))
        # K/V caches addressed through a beam indirection table by unimp_lm_decode_attn (default), or
        # SDPA over caches that are re-ordered (copied) after every beam step (UNIMP_DECODE_INDIR=0)
        self.use_indirection = os.environ.get("UNIMP_DECODE_INDIR", "1") != "0"
        self.last_n_steps, self.last_logits_finite = 0, True   # of the last generate() (diagnostics)

    # -------------------------------------------------------------------------------- one token
    def _token_step(self, S):
        """tokens (Bf,1) -> logits (Bf,V); writes K/V at S.cur.  Everything here is captured."""
        from .flamingo_lm import fused_neox_layer

        lm = self.lm
        x = lm.get_input_embeddings()(S.tokens)                              # (Bf,1,D)
        pos_emb = lm.gpt_neox.rotary_emb(x, position_ids=S.pos[:, None])
        # the new key position becomes visible; everything beyond stays masked
        S.add_mask.index_fill_(3, S.cur, 0.0)
        # every LayerNorm rides on the residual add that produces its input (one K5 launch): the block /
        # layer that ends at x also returns first_ln_of_the_next(x), as FlamingoLayer.forward does
        x_ln = None
        n_layers = len(self.layers)
        for i, layer in enumerate(self.layers):
            blk, dl = layer.gated_cross_attn_layer, layer.decoder_layer
            next_ln = self.layers[i + 1].first_ln() if i + 1 < n_layers else lm.gpt_neox.final_layer_norm
            h1 = x_ln
            if blk is not None:
                x, h1 = blk(x, layer.vis_x, media_locations=layer.media_locations, use_cached_media=True,
                            text_time=S.n_media, next_ln=dl.input_layernorm, x_ln=x_ln)
            kv_step = ((S.k[i], S.v[i], S.cur, S.indir, S.mask2d) if S.indir is not None
                       else (S.k[i], S.v[i], S.cur))
            x, x_ln = fused_neox_layer(dl, x, S.add_mask, pos_emb, h1=h1, next_ln=next_ln, kv_step=kv_step)
        S.logits.copy_(lm.embed_out(x_ln)[:, -1, :])       # x_ln == final_layer_norm(x)
        S.cur += 1
        S.pos += 1

    @staticmethod
    def _capture(fn):
        """Capture `fn` (one decode step over static buffers) into a CUDA graph; nothing runs yet."""
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        return graph

    def _reorder(self, S, source):
        """New beam i continues old beam source[i].  With the indirection table only ITS rows move
        (B*beams x T_max ints); the caches stay where they are and `unimp_lm_decode_attn` reads
        position t of beam i from cache row indir[i][t].  (UNIMP_DECODE_INDIR=0: HF's `reorder_cache`
        semantics, every layer's K and V gathered — 2 x 32 x 20 MB per token at configs[3].)"""
        if S.indir is not None:
            S.indir.copy_(S.indir.index_select(0, source))
            return
        for i in range(len(self.layers)):
            S.k[i].copy_(S.k[i].index_select(0, source))
            S.v[i].copy_(S.v[i].index_select(0, source))

    # -------------------------------------------------------------------------------- generate
    @torch.no_grad()
    def generate(self, vision_x, lang_x, attention_mask=None, *, num_beams: int = 1, max_new_tokens: int = 20,
                 eos_token_id=None, pad_token_id=None, num_return_sequences: int = 1, early_stopping=False,
                 length_penalty: float = 1.0, do_sample: bool = False, no_repeat_ngram_size: int = 0,
                 **unsupported):
        """Arguments as `Flamingo.generate` / HF `generate` (reference call:
        `UniMP/pipeline/eval/eval_exp.py:101-114`).  Deterministic decoding only: sampling, n-gram
        blocking and other logits processors raise — use `Flamingo.generate` for those."""
        from types import SimpleNamespace

        if do_sample or no_repeat_ngram_size or unsupported:
            raise NotImplementedError(
                "GraphedDecoder implements greedy / beam search without logits processors; got "
                f"do_sample={do_sample}, no_repeat_ngram_size={no_repeat_ngram_size}, "
                f"{sorted(unsupported)} — use Flamingo.generate for these")
        if num_return_sequences > num_beams:
            raise ValueError("num_return_sequences has to be smaller or equal to num_beams")
        model, lm = self.model, self.lm
        assert lang_x.is_cuda, "GraphedDecoder needs a CUDA device (there is no CPU path)"
        eos = model.eoc_token_id if eos_token_id is None else eos_token_id
        B, T0 = lang_x.shape
        nb = num_beams
        Bf = B * nb
        Tmax = (T0 + max_new_tokens + 15) // 16 * 16     # cache length; the tail past the cursor stays masked
        dev = lang_x.device
        ids = lang_x.repeat_interleave(nb, dim=0)
        mask = torch.ones_like(ids) if attention_mask is None else attention_mask.repeat_interleave(nb, dim=0)
        mask = mask.to(torch.int64)

        timed = torch.cuda.is_available() and lang_x.device.type == "cuda"   # device-side phase durations
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None

        def stamp(i):
            if timed:
                ev[i].record()

        stamp(0)
        # ---- prefill: the ordinary module path, vision latents cached on the layers --------------
        was_training = model.training
        model.eval()
        lm._use_cached_vision_x = True
        model._encode_vision_x(vision_x=vision_x.repeat_interleave(nb, dim=0) if nb > 1 else vision_x)
        pos_ids = (mask.cumsum(-1) - 1).masked_fill(mask == 0, 1)
        out = lm(input_ids=ids, attention_mask=mask, position_ids=pos_ids, use_cache=True, logits_to_keep=1)
        cache = out.past_key_values
        dtype = out.logits.dtype

        # ---- static state ----------------------------------------------------------------------
        S = SimpleNamespace()
        S.tokens = torch.zeros((Bf, 1), dtype=torch.int64, device=dev)
        S.pos = mask.sum(-1)                                             # rotary position of the next token
        S.cur = torch.full((1,), T0, dtype=torch.int64, device=dev)      # K/V slot of the next token
        S.n_media = (ids == lm.media_token_id).sum(-1, keepdim=True).to(torch.int32)
        S.logits = torch.zeros((Bf, out.logits.shape[-1]), dtype=torch.float32, device=dev)
        S.add_mask = torch.full((Bf, 1, 1, Tmax), self.MASK_FILL, dtype=dtype, device=dev)
        S.add_mask[:, 0, 0, :T0] = torch.zeros((), dtype=dtype, device=dev).expand(Bf, T0).masked_fill(mask == 0, self.MASK_FILL)
        # beam indirection: cache row that holds position t of beam i's history (identity at first)
        S.indir = S.mask2d = None
        head_dim = self.layers[0].decoder_layer.attention.head_size
        # unimp_lm_decode_attn covers head dims 32..128 and caches up to 5120 positions; beyond that the
        # step runs on re-ordered caches + SDPA (same tokens, tests/test_decode.py)
        if self.use_indirection and Tmax <= 5120 and head_dim in (32, 64, 80, 96, 128):
            S.indir = torch.arange(Bf, dtype=torch.int32, device=dev)[:, None].repeat(1, Tmax).contiguous()
            S.mask2d = S.add_mask.view(Bf, Tmax)
        S.k, S.v = [], []
        for i in range(len(self.layers)):
            k, v = cache.layers[i].keys, cache.layers[i].values          # (Bf,H,T0,dh)
            kb = torch.zeros((Bf, k.shape[1], Tmax, k.shape[3]), dtype=k.dtype, device=dev)
            vb = torch.zeros_like(kb)
            kb[:, :, :T0], vb[:, :, :T0] = k, v
            S.k.append(kb)
            S.v.append(vb)
        del cache
        for layer in self.layers:                                        # to_kv(media): once per block
            blk = layer.gated_cross_attn_layer
            if blk is not None:
                blk.attn.cached_media_kv(layer.vis_x)
        first_logits = out.logits[:, -1, :].float()
        del out

        search = make_search(lang_x, num_beams=nb, max_new_tokens=max_new_tokens, eos_token_id=eos,
                             pad_token_id=pad_token_id, length_penalty=length_penalty,
                             early_stopping=early_stopping)

        def advance(logits):
            tok, source = search.step(logits)
            if source is not None:
                self._reorder(S, source)
            S.tokens.copy_(tok[:, None])

        def one_token():
            self._token_step(S)
            advance(S.logits)

        # step 0 consumes the prefill logits (eagerly); every later token is a graph replay
        advance(first_logits)
        stamp(1)
        n_steps = 1
        graph = None
        try:
            if n_steps < max_new_tokens and bool(search.unfinished()):
                one_token()                                              # eager warm-up step
                n_steps += 1
            if n_steps < max_new_tokens and bool(search.unfinished()):
                graph = self._capture(one_token)
            stamp(2)
            n_replay = 0
            while graph is not None and n_steps < max_new_tokens:
                if n_steps % self.sync_every == 0 and not bool(search.unfinished()):
                    break
                graph.replay()
                n_steps += 1
                n_replay += 1
            stamp(3)
            result = search.result(num_return_sequences)                 # (reads device state: syncs)
            if timed:
                torch.cuda.synchronize()
                self.last_timing = {"prefill_ms": ev[0].elapsed_time(ev[1]),
                                    "eager_step_and_capture_ms": ev[1].elapsed_time(ev[2]),
                                    "replay_ms": ev[2].elapsed_time(ev[3]), "replays": n_replay,
                                    "tokens": n_steps}
            self.last_n_steps = n_steps
            self.last_logits_finite = bool(torch.isfinite(S.logits).all())
        finally:
            del graph
            lm.clear_conditioned_layers()
            lm._use_cached_vision_x = False
            for layer in self.layers:
                if layer.gated_cross_attn_layer is not None:
                    layer.gated_cross_attn_layer.attn._kv_cache = None
            model.train(was_training)
        return result
