"""FlamingoLayer / FlamingoLMMixin — the language-model side of the drop-in.

Same names, attributes and conditioning protocol as `open_flamingo/src/flamingo_lm.py`
v2.0.1: the mixin is grafted onto a HF causal LM *instance*, builds
`gated_cross_attn_layers`, wraps each decoder layer in a FlamingoLayer, and its forward
derives media locations from `input_ids == media_token_id`.  New here: the per-token image
index (`text_time`, int32) is computed ONCE per forward by the `unimp_text_time` kernel and
handed to every layer, instead of a bool mask being re-cumsum'ed and materialised per layer;
and `forward(labels=...)` gets its logged mean CE (`out[0]`, reference UniMP/mmrec.py:182)
from the one-pass focal-CE kernel rather than HF's fp32 upcast + CrossEntropyLoss.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
from transformers.modeling_outputs import CausalLMOutputWithPast

from . import ops
from .helpers import GatedCrossAttentionBlock


@dataclass
class LabelRowsOutput(CausalLMOutputWithPast):
    """`forward(..., label_rows=...)`: `logits` holds ONLY the rows the loss reads — (R, V), row r
    = position `row_index[r]` of the flattened (B*T) sequence, scored against `row_targets[r]`
    (-100 = unused slot of a fixed-capacity gather).  `loss` (= out[0]) is the HF mean CE."""
    row_index: Optional[torch.LongTensor] = None
    row_targets: Optional[torch.LongTensor] = None
    overflow: Optional[torch.Tensor] = None


def getattr_recursive(obj, att):
    if att == "":
        return obj
    i = att.find(".")
    if i < 0:
        return getattr(obj, att)
    return getattr_recursive(getattr(obj, att[:i]), att[i + 1:])


def setattr_recursive(obj, att, val):
    if "." in att:
        obj = getattr_recursive(obj, ".".join(att.split(".")[:-1]))
    setattr(obj, att.split(".")[-1], val)


def extend_instance(obj, mixin):
    """Apply a mixin to an existing instance (upstream `utils.extend_instance`)."""
    base_cls = obj.__class__
    obj.__class__ = type(base_cls.__name__, (mixin, base_cls), {})


class _PaddedHeadFn(torch.autograd.Function):
    """logits = h @ W^T with W zero-padded to a multiple of 128 rows, returned as the [:V] view.

    V = 74 053 is odd: an unaligned N sends cuBLAS to a legacy kernel (measured 2.5 ms vs 0.4 ms
    per call on B200).  The pad columns never leave this function: the focal-CE kernels take the
    padded row stride as `ld`, and their d_logits comes back as the [:V] view of a buffer with the
    same padded stride, which backward re-expands in place (no copy) for dH = dlogits @ W."""

    @staticmethod
    def forward(ctx, h, w_pad, V):
        ctx.save_for_backward(w_pad)
        ctx.V = V
        return torch.nn.functional.linear(h, w_pad)[..., :V]

    @staticmethod
    def backward(ctx, g):
        (w_pad,) = ctx.saved_tensors
        V, Vp = ctx.V, w_pad.shape[0]
        lead = g.shape[:-1]
        n_lead = 1
        for d in lead:
            n_lead *= d
        need = (g.storage_offset() + n_lead * Vp) * g.element_size()
        dense_prefix = g.dim() >= 2 and g.stride(-1) == 1 and g.stride(-2) == Vp and all(
            g.stride(i) == g.stride(i + 1) * g.shape[i + 1] for i in range(g.dim() - 2))
        if dense_prefix and g.untyped_storage().nbytes() >= need:
            gp = g.as_strided(tuple(lead) + (Vp,), tuple(g.stride()[:-1]) + (1,), g.storage_offset())
            gp[..., V:].zero_()
        else:
            gp = torch.nn.functional.pad(g, (0, Vp - V))
        return (gp.reshape(-1, Vp) @ w_pad).reshape(*lead, w_pad.shape[1]), None, None


class PaddedOutputHead(nn.Linear):
    """Drop-in for the frozen, bias-free `embed_out` Linear: still an `nn.Linear` (state-dict key
    `weight`, shape (V, D), `in_features` / `out_features` unchanged), so everything that inspects
    the head — HF `resize_token_embeddings` (reference `UniMP/mmrec.py:595`), `tie_weights`,
    checkpoint code — keeps working; see _PaddedHeadFn for what forward does differently."""

    def __init__(self, linear: nn.Linear, multiple: int = 128):
        assert linear.bias is None
        nn.Module.__init__(self)
        self.in_features, self.out_features = linear.in_features, linear.out_features
        self.weight = linear.weight
        self.register_parameter("bias", None)
        self.multiple = multiple
        self._wp = None

    def _padded(self):
        w = self.weight
        V = w.shape[0]
        Vp = (V + self.multiple - 1) // self.multiple * self.multiple
        wp = self._wp
        if wp is None or wp.dtype != w.dtype or wp.device != w.device or wp.data_ptr() != w.data_ptr():
            wp = torch.zeros((Vp, w.shape[1]), dtype=w.dtype, device=w.device)
            wp[:V].copy_(w.data)
            w.data = wp[:V]          # the parameter now aliases the padded buffer: no second copy
            self._wp = wp
        return wp

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        # the parameter may be a view of the padded buffer: serialise exactly (V, D), own storage
        w = self.weight if keep_vars else self.weight.detach().clone()
        destination[prefix + "weight"] = w

    def forward(self, h):
        if self.weight.requires_grad or not h.is_cuda:
            return torch.nn.functional.linear(h, self.weight)
        return _PaddedHeadFn.apply(h, self._padded(), self.out_features)

    def gathered(self, h_rows):
        """logits of a (N, D) row selection, padded stride, [:V] view (head+loss fusion)."""
        return _PaddedHeadFn.apply(h_rows, self._padded(), self.out_features)


def wrap_output_head(lm):
    """(Re-)wrap `lm.embed_out` when its width needs padding for the aligned cuBLAS path."""
    head = lm.get_output_embeddings()
    if isinstance(head, PaddedOutputHead):
        return head
    if (isinstance(head, nn.Linear) and head.bias is None and not head.weight.requires_grad
            and head.out_features % 8):
        lm.set_output_embeddings(PaddedOutputHead(head))
    return lm.get_output_embeddings()


def _neox_fusable(layer, x, kw) -> bool:
    """Can this (frozen) HF GPT-NeoX decoder layer run on the fused path? (no KV cache, SDPA,
    no active dropout, vectorisable rotary dims)"""
    from transformers.models.gpt_neox.modeling_gpt_neox import GPTNeoXLayer

    if not isinstance(layer, GPTNeoXLayer) or not x.is_cuda:
        return False
    if x.dtype not in (torch.bfloat16, torch.float32):
        return False
    if kw.get("layer_past") is not None or kw.get("position_embeddings") is None:
        return False
    att = layer.attention
    if getattr(att.config, "_attn_implementation", "sdpa") != "sdpa":
        return False
    if layer.training and (layer.post_attention_dropout.p > 0 or layer.post_mlp_dropout.p > 0
                           or att.attention_dropout > 0):
        return False
    npv = 8 if x.dtype == torch.bfloat16 else 4
    return att.rotary_ndims % 2 == 0 and (att.rotary_ndims // 2) % npv == 0 and att.head_size % npv == 0


def _is_exact_gelu(act) -> bool:
    """GPT-NeoX `hidden_act="gelu"` (RedPajama-INCITE): HF GELUActivation / nn.GELU, erf form."""
    from transformers.activations import GELUActivation

    if isinstance(act, GELUActivation):
        return act.act is torch.nn.functional.gelu
    return isinstance(act, torch.nn.GELU) and act.approximate == "none"


# K4 on our own kernel (bf16, head dim 80); UNIMP_LM_ATTN=0 keeps cuDNN SDPA for A/B runs
LM_ATTN = os.environ.get("UNIMP_LM_ATTN", "1") != "0"

_MASK_CACHE = [None, None, None]   # (bool mask object, (dtype, version), additive mask)


def _additive_mask(mask, dtype):
    """SDPA turns a boolean attn_mask into an additive one on EVERY call (one `where` launch per
    decoder layer).  HF hands the same mask tensor to all layers of a forward: convert it once.
    The cache holds a reference to the mask it was built from (identity, not address, is the key)."""
    if mask is None or mask.dtype != torch.bool:
        return mask
    key = (dtype, mask._version)       # an in-place edit of the same mask object invalidates the entry
    if _MASK_CACHE[0] is not mask or _MASK_CACHE[1] != key:
        add = torch.zeros(mask.shape, dtype=dtype, device=mask.device).masked_fill_(~mask, float("-inf"))
        _MASK_CACHE[0], _MASK_CACHE[1], _MASK_CACHE[2] = mask, key, add
    return _MASK_CACHE[2]


def fused_neox_layer(layer, x, attention_mask, position_embeddings, h1=None, next_ln=None, kv_step=None,
                     key_bits=None):
    """HF `GPTNeoXLayer.forward` (transformers gpt_neox, no cache) with its elementwise glue on
    our kernels: LayerNorms and residual adds are K5 launches, rotary runs on the packed qkv
    projection (`unimp_rotary_qkv_*`), the causal attention core is `unimp_lm_attn_*` (K4: bf16,
    head dim 80, `key_bits` given) and otherwise SDPA over strided views.  GEMMs stay on cuBLAS.
    Same parameters, same arithmetic; ~11 launches instead of ~30 per layer.
    `key_bits`: True = causal only, or `ops.key_bits(attention_mask_2d)`; the caller vouches that
    `attention_mask` (the 4-D mask HF built) is exactly causal & those key bits.
    `h1`: input_layernorm(x) if the caller already produced it in a fused epilogue.
    `next_ln`: the LayerNorm that reads this layer's output first (next layer's x-attn norm or
    input_layernorm); if given, returns (y, next_ln(y)) from the same launch as the last residual.
    `kv_step` = (k_cache, v_cache, cursor): single-token decode against static (B,H,T_max,dh)
    caches — the new key/value are written at the device-side `cursor` (capturable in a CUDA
    graph) and attention runs over the whole cache under the additive `attention_mask` (SDPA).
    `kv_step` = (k_cache, v_cache, cursor, indir, mask2d): the same step on `unimp_lm_decode_attn`
    (beam indirection `indir` (B,T_max) int32 instead of re-ordered caches, `mask2d` (B,T_max) 0/-inf)."""
    F = torch.nn.functional
    att = layer.attention
    B, T, D = x.shape
    H, dh, rot = att.config.num_attention_heads, att.head_size, att.rotary_ndims
    ln1, ln2 = layer.input_layernorm, layer.post_attention_layernorm
    if h1 is None:
        h1 = ops.layer_norm(x, ln1.weight, ln1.bias, ln1.eps)
    # decode steps (<= 8 rows, no autograd) stream each weight once through unimp_linear_small_m
    lin = ops.linear_rows if kv_step is not None else (lambda a, w, b, act_gelu=False: F.linear(a, w, b))
    qkv = lin(h1, att.query_key_value.weight, att.query_key_value.bias)
    cos, sin = position_embeddings
    if kv_step is not None and len(kv_step) == 5:
        # one new token against static caches with beam indirection (unimp_lm_decode_attn): rotary,
        # the cache write at the cursor and the attention in ONE launch; beams are never copied
        k_cache, v_cache, cursor, indir, mask2d = kv_step
        assert T == 1
        a = ops.lm_decode_attention(qkv, cos.to(x.dtype), sin.to(x.dtype), k_cache, v_cache, indir, mask2d,
                                    cursor, heads=H, head_dim=dh, rotary_dim=rot, scale=att.scaling)
    elif key_bits is not None and kv_step is None and LM_ATTN and ops.lm_attention_supported_shape(x, T, H, dh):
        a = ops.rotary_lm_attention(qkv, cos.to(x.dtype), sin.to(x.dtype), None if key_bits is True else key_bits,
                                    heads=H, head_dim=dh, rotary_dim=rot, scale=att.scaling)
    else:
        q, k, v = ops.rotary_qkv(qkv, cos.to(x.dtype), sin.to(x.dtype), heads=H, head_dim=dh,
                                 rotary_dim=rot)
        attention_mask = _additive_mask(attention_mask, x.dtype)
        if kv_step is not None:
            k_cache, v_cache, cursor = kv_step
            k_cache.index_copy_(2, cursor, k)
            v_cache.index_copy_(2, cursor, v)
            k, v = k_cache, v_cache
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0,
                                           is_causal=attention_mask is None and T > 1 and kv_step is None,
                                           scale=att.scaling)
        a = a.transpose(1, 2).reshape(B, T, D)
    o = lin(a, att.dense.weight, att.dense.bias)
    mlp = layer.mlp
    exact = _is_exact_gelu(mlp.act)
    act = ops.gelu if exact else mlp.act

    def mlp_fwd(h2):
        if kv_step is not None and exact:      # GELU in the epilogue of the first projection
            u = lin(h2, mlp.dense_h_to_4h.weight, mlp.dense_h_to_4h.bias, act_gelu=True)
        else:
            u = act(lin(h2, mlp.dense_h_to_4h.weight, mlp.dense_h_to_4h.bias))
        return lin(u, mlp.dense_4h_to_h.weight, mlp.dense_4h_to_h.bias)

    if layer.use_parallel_residual:
        h2 = ops.layer_norm(x, ln2.weight, ln2.bias, ln2.eps)
        m = mlp_fwd(h2)
        x1 = ops.gate_residual(o, x, None)
    else:
        x1, h2 = ops.gate_residual_ln(o, x, None, ln2.weight, ln2.bias, ln2.eps)
        m = mlp_fwd(h2)
    if next_ln is not None:
        return ops.gate_residual_ln(m, x1, None, next_ln.weight, next_ln.bias, next_ln.eps)
    return ops.gate_residual(m, x1, None)


class FlamingoLayer(nn.Module):
    def __init__(self, gated_cross_attn_layer, decoder_layer, gradient_checkpointing=False):
        super().__init__()
        self.gated_cross_attn_layer = gated_cross_attn_layer
        self.decoder_layer = decoder_layer
        self.vis_x = None
        self.media_locations = None
        self.text_time = None
        self.use_cached_media = False
        self.lm_key_bits = None      # set per forward by FlamingoLMMixin.forward (fused_neox_layer)
        self._next_holder = [None]   # next FlamingoLayer (a list: not a registered submodule)
        self._pre = None             # (x, first_ln(x)) handed over by the previous layer

    def first_ln(self):
        """The LayerNorm that reads this layer's input first."""
        if self.gated_cross_attn_layer is not None:
            return self.gated_cross_attn_layer.attn.norm
        return getattr(self.decoder_layer, "input_layernorm", None)

    def is_conditioned(self) -> bool:
        return self.vis_x is not None and self.media_locations is not None

    def condition_vis_x(self, vis_x):
        self.vis_x = vis_x
        if self.gated_cross_attn_layer is not None:
            self.gated_cross_attn_layer.attn._kv_cache = None   # new media (or none): never stale

    def condition_media_locations(self, media_locations, text_time=None):
        self.media_locations = media_locations
        self.text_time = text_time

    def condition_use_cached_media(self, use_cached_media):
        self.use_cached_media = use_cached_media

    def forward(self, lang_x, attention_mask=None, **decoder_layer_kwargs):
        fuse = _neox_fusable(self.decoder_layer, lang_x, decoder_layer_kwargs)
        pre, self._pre = self._pre, None
        x_ln = pre[1] if (pre is not None and pre[0] is lang_x) else None  # first_ln(lang_x), fused upstream
        h1 = None if self.gated_cross_attn_layer is not None else x_ln
        nxt = self._next_holder[0]
        next_ln = nxt.first_ln() if (fuse and nxt is not None) else None
        if self.gated_cross_attn_layer is not None:
            if self.vis_x is None:
                raise ValueError("vis_x must be conditioned before forward pass")
            if self.media_locations is None:
                raise ValueError("media_locations must be conditioned before forward pass")
            tt = self.text_time
            if tt is not None and tt.shape[1] != lang_x.shape[1]:
                tt = None  # stale (e.g. cached prompt length): let the block derive it
            out = self.gated_cross_attn_layer(
                lang_x, self.vis_x, media_locations=self.media_locations,
                use_cached_media=self.use_cached_media, text_time=tt,
                next_ln=self.decoder_layer.input_layernorm if fuse else None, x_ln=x_ln)
            lang_x, h1 = out if fuse else (out, None)
        if fuse:
            out = fused_neox_layer(self.decoder_layer, lang_x, attention_mask,
                                   decoder_layer_kwargs["position_embeddings"], h1, next_ln,
                                   key_bits=self.lm_key_bits)
            if next_ln is not None:
                y, y_ln = out
                nxt._pre = (y, y_ln)   # consumed (and cleared) by the next layer's forward
                return y
            return out
        return self.decoder_layer(lang_x, attention_mask=attention_mask, **decoder_layer_kwargs)


class FlamingoLMMixin(nn.Module):
    """Mixin adding gated cross-attention to a HF causal LM instance."""

    def set_decoder_layers_attr_name(self, name):
        self.decoder_layers_attr_name = name

    def _get_decoder_layers(self):
        return getattr_recursive(self, self.decoder_layers_attr_name)

    def _set_decoder_layers(self, value):
        setattr_recursive(self, self.decoder_layers_attr_name, value)

    def init_flamingo(self, media_token_id, lang_hidden_size, vis_hidden_size,
                      cross_attn_every_n_layers, gradient_checkpointing=False):
        self.old_decoder_blocks = self._get_decoder_layers()
        self.gated_cross_attn_layers = nn.ModuleList([
            GatedCrossAttentionBlock(dim=lang_hidden_size, dim_visual=vis_hidden_size)
            if (i + 1) % cross_attn_every_n_layers == 0 else None
            for i, _ in enumerate(self._get_decoder_layers())
        ])
        self.init_flamingo_layers(gradient_checkpointing)
        self.media_token_id = media_token_id
        self.initialized_flamingo = True
        self._use_cached_vision_x = False

    def init_flamingo_layers(self, gradient_checkpointing=False):
        layers = [FlamingoLayer(g, d, gradient_checkpointing)
                  for g, d in zip(self.gated_cross_attn_layers, self.old_decoder_blocks)]
        for a, b in zip(layers[:-1], layers[1:]):
            a._next_holder[0] = b
        self._set_decoder_layers(nn.ModuleList(layers))

    def forward(self, input_ids=None, attention_mask=None, labels=None, label_rows=None, **kwargs):
        """`label_rows` (product-only, default off = upstream behaviour): head + loss fusion for
        training.  The reference's loss (UniMP/mmrec.py:190-213) and HF's logged
        mean CE read only rows whose shifted label is not -100 (24 of 1536 at configs[1]); with
        `label_rows=True` (exact, one host sync) or an int capacity (static shapes, graph-safe)
        those hidden rows are gathered BEFORE `embed_out`, so the head GEMM, the logits and
        d_logits are (R, V) instead of (B, T, V).  Returns a LabelRowsOutput."""
        if not self.initialized_flamingo:
            raise ValueError("Flamingo layers are not initialized. Call `init_flamingo` first.")
        media_locations = input_ids == self.media_token_id
        # single host sync per forward only on the decode path, exactly where upstream has it
        use_cached = bool(self._use_cached_vision_x and self.is_conditioned()
                          and not media_locations.any())
        layers = self._get_decoder_layers()
        if use_cached:
            cached = layers[0].media_locations
            tt = ops.text_time(cached.to(torch.int64), 1, use_cached=True,
                               T_out=input_ids.shape[1])
            for layer in layers:
                layer.text_time = tt
                layer.condition_use_cached_media(True)
        else:
            tt = ops.text_time(input_ids, self.media_token_id)
            for layer in layers:
                layer.condition_media_locations(media_locations, tt)
                layer.condition_use_cached_media(False)
        emb = self.get_input_embeddings()
        if (torch.is_grad_enabled() and emb.weight.requires_grad and input_ids.is_cuda
                and kwargs.get("inputs_embeds") is None and type(emb) is nn.Embedding
                and emb.padding_idx is None and emb.max_norm is None):
            # trainable input embeddings: gradient rows go straight into the flat grad buffer
            kwargs["inputs_embeds"] = ops.embedding_acc(input_ids, emb.weight)
        else:
            kwargs["input_ids"] = input_ids
        kwargs["attention_mask"] = attention_mask
        # K4: with no KV cache HF's 4-D mask is exactly causal & key padding; hand the layers the
        # padding as one bit per key (they fall back to SDPA with HF's mask for anything else)
        kb = None
        if input_ids.is_cuda and kwargs.get("past_key_values") is None and not use_cached:
            if attention_mask is None:
                kb = True
            elif attention_mask.dim() == 2 and attention_mask.shape == input_ids.shape:
                kb = ops.key_bits(attention_mask)
        for layer in layers:
            layer.lm_key_bits = kb
        try:
            if label_rows is not None and label_rows is not False:
                return self._forward_label_rows(labels, label_rows, kwargs)
            out = super().forward(**kwargs)  # HF forward without labels: no fp32 logits copy
        finally:
            for layer in layers:
                layer.lm_key_bits = None
        if labels is None:
            return out
        logits = out.logits
        ones = torch.ones(logits.shape[0], dtype=torch.float32, device=logits.device)
        loss = ops.focal_ce(logits, labels, ones, gamma=0.0, use_focal=False)  # HF mean CE
        return CausalLMOutputWithPast(loss=loss, logits=logits,
                                      past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states, attentions=out.attentions)

    def _forward_label_rows(self, labels, label_rows, kwargs):
        if labels is None:
            raise ValueError("label_rows needs labels (it gathers the rows the loss reads)")
        kwargs.pop("logits_to_keep", None)
        base = self.base_model(**kwargs)                       # decoder stack + final LayerNorm
        hidden = base.last_hidden_state                        # (B, T, D)
        B, T, D = hidden.shape
        cap = None if label_rows is True else int(label_rows)
        idx, targets, overflow = ops.gather_label_rows(labels.to(hidden.device), cap)
        rows = hidden.reshape(B * T, D).index_select(0, idx)   # backward: scatter-add of R rows
        head = self.get_output_embeddings()
        logits = head.gathered(rows) if isinstance(head, PaddedOutputHead) and not head.weight.requires_grad \
            and rows.is_cuda else head(rows)
        ones = torch.ones(idx.shape[0], dtype=torch.float32, device=logits.device)
        loss = ops.focal_ce_rows(logits, targets, ones, None, n_groups=1, gamma=0.0, use_focal=False)
        return LabelRowsOutput(loss=loss, logits=logits, past_key_values=base.past_key_values,
                               hidden_states=base.hidden_states, attentions=base.attentions,
                               row_index=idx, row_targets=targets, overflow=overflow)

    def resize_token_embeddings(self, new_num_tokens=None, *args, **kwargs):
        """reference `UniMP/mmrec.py:595` (`lang_encoder.resize_token_embeddings(len(tokenizer))`
        after the vocabulary grew): unwrap the padded head to a plain nn.Linear of exactly (V, D),
        let HF resize both embeddings, re-wrap for the new width."""
        head = self.get_output_embeddings()
        if isinstance(head, PaddedOutputHead):
            plain = nn.Linear(head.in_features, head.out_features, bias=False,
                              device=head.weight.device, dtype=head.weight.dtype)
            with torch.no_grad():
                plain.weight.copy_(head.weight)
            plain.weight.requires_grad_(head.weight.requires_grad)
            self.set_output_embeddings(plain)
        out = super().resize_token_embeddings(new_num_tokens, *args, **kwargs)
        new_head = self.get_output_embeddings()
        if isinstance(head, PaddedOutputHead) and new_head is not None:
            new_head.weight.requires_grad_(head.weight.requires_grad)
        wrap_output_head(self)
        return out

    def is_conditioned(self) -> bool:
        return all(l.is_conditioned() for l in self._get_decoder_layers())

    def clear_conditioned_layers(self):
        for layer in self._get_decoder_layers():
            layer.condition_vis_x(None)
            layer.condition_media_locations(None)
            layer.condition_use_cached_media(None)
