"""Restatement of UniMP's in-tree label masking and focal-loss head (oracle).

TEST INFRASTRUCTURE — see oracle/__init__.py.  PINNED: this arithmetic IS in the reference
tree, every function cites the lines it follows, and tests/test_reference_golden.py checks it
(labels bit-exact, loss and dloss/dlogits to 1e-6) against vectors obtained by running the
unmodified `UniMP/mmrec.py::train_one_epoch` (tests/golden/make_reference_golden.py).
"""
from __future__ import annotations

import torch


def mask_labels(input_ids, *, answer_token_id, endofchunk_token_id, media_token_id,
                pad_token_id):
    """Answer-span label masking; follows reference `UniMP/mmrec.py:143-168` line by line.

    Keeps tokens strictly after `<answer>` up to (excluding) `<|endofchunk|>`; drops pad,
    position 0, `<answer>` and `<image>`.  The trailing EOS after the last answer is kept
    (the state machine never leaves the answer state).  Pure-Python O(B*T) loop on purpose.
    """
    labels = input_ids.clone()
    for i in range(labels.shape[0]):            # mmrec.py:146
        answer_flag = 0                         # mmrec.py:147
        for j in range(labels.shape[1]):        # mmrec.py:148
            if not answer_flag:                 # mmrec.py:149
                if labels[i, j] == answer_token_id:
                    answer_flag = 1             # mmrec.py:150-151
                labels[i, j] = -100             # mmrec.py:152
            else:
                if labels[i, j] == endofchunk_token_id:   # mmrec.py:154
                    answer_flag = 0
                    labels[i, j] = -100         # mmrec.py:155-156
    labels[labels == pad_token_id] = -100       # mmrec.py:157
    labels[:, 0] = -100                         # mmrec.py:158
    labels[labels == answer_token_id] = -100    # mmrec.py:167
    labels[labels == media_token_id] = -100     # mmrec.py:168
    return labels


def focal_loss(lm_logits, labels, weights, *, gamma=2.0, use_reweight=True):
    """Task-weighted focal CE; follows reference `UniMP/mmrec.py:190-213`.

    lm_logits (B,T,V), labels (B,T) with -100, weights (B,).  Returns the scalar
    `sum(w * CE * (1-pt)^gamma) / sum(labels != -100)` on shift-by-one logits/labels.
    Quirks kept: `p[rows, labels]` wraps -100 to column V-100 (harmless: CE is 0 there,
    needs V >= 100); the focal factor is NOT detached; an all-ignored batch gives NaN.
    """
    labels = labels.to(lm_logits.device)
    n1, n2 = labels.shape[0], labels.shape[1] - 1              # mmrec.py:193
    shift_logits = lm_logits[:, :-1, :].contiguous()            # mmrec.py:194
    labels = labels[:, 1:].contiguous()                         # mmrec.py:195
    loss_fct = torch.nn.CrossEntropyLoss(reduction="none")      # mmrec.py:196
    shift_logits = shift_logits.view(-1, shift_logits.size(-1)) # mmrec.py:198
    labels = labels.view(-1)                                    # mmrec.py:199
    lm_loss = loss_fct(shift_logits, labels).view(n1, n2)       # mmrec.py:201
    loss = torch.unsqueeze(weights, 1) * lm_loss                # mmrec.py:203
    loss = loss.view(-1)                                        # mmrec.py:204
    if use_reweight:                                            # mmrec.py:205
        p = torch.nn.functional.softmax(shift_logits, dim=-1)   # mmrec.py:207
        all_rows = torch.arange(len(shift_logits))              # mmrec.py:208
        pt = p[all_rows, labels]                                # mmrec.py:209
        focal_term = (1 - pt) ** gamma                          # mmrec.py:210
        loss = loss * focal_term                                # mmrec.py:212
    return torch.sum(loss) / torch.sum(labels != -100)          # mmrec.py:213


def focal_loss_closed_form_grad(lm_logits, labels, weights, *, gamma=2.0, use_reweight=True):
    """d loss / d lm_logits in closed form; oracle self-check and the
    formula the CUDA backward implements:
    (w/N_valid) * [(1-pt)^g + g*(1-pt)^(g-1)*pt*CE] * (p - onehot), 0 on ignored rows
    and on the last time step."""
    B, T, V = lm_logits.shape
    z = lm_logits[:, :-1, :].double()
    y = labels[:, 1:]
    valid = y != -100
    nvalid = valid.sum()
    p = torch.softmax(z, dim=-1)
    ysafe = y.clamp(min=0)
    pt = p.gather(-1, ysafe[..., None]).squeeze(-1)
    ce = -torch.log(pt)
    if use_reweight:
        coef = (1 - pt) ** gamma + gamma * (1 - pt) ** (gamma - 1) * pt * ce
    else:
        coef = torch.ones_like(pt)
    coef = coef * weights[:, None].double() / nvalid
    coef = torch.where(valid, coef, torch.zeros_like(coef))
    onehot = torch.zeros_like(p).scatter_(-1, ysafe[..., None], 1.0)
    g = coef[..., None] * (p - onehot)
    out = torch.zeros(B, T, V, dtype=torch.float64)
    out[:, :-1] = g
    return out
