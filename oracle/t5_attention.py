"""TEST INFRASTRUCTURE - CPU restatement of T5Attention.forward (genrec/modules/transformer.py:13-159) as plain functions over a
``state_dict``; only tests/ may import this.  Pinned by tests/golden/t5_attention.pt (outputs and gradients of the UNMODIFIED
reference module, oracle/make_golden.py)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def bucket_of(rel: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """Bidirectional T5 bucket of rel = memory_position - context_position (transformer.py:13-41)."""
    n = -rel
    nb = num_buckets // 2
    sign = (n < 0).long()
    n = n.abs()
    max_exact = nb // 2
    log_part = (torch.log(n.float() / max_exact + 1e-6) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = max_exact + log_part.clamp(max=nb - max_exact - 1)          # the clamp applies to the logarithmic part only (:31-35)
    return torch.where(n < max_exact, n, large) + sign * nb


def t5_attention_forward(query, key, value, sd, n_heads, is_cross, attn_mask=None, key_padding_mask=None, num_buckets=32, max_distance=128,
                         prefix=""):
    """sd: {"q.weight", "kv.weight" | "k.weight" + "v.weight", "o.weight", "rel_bias.weight"?} -> out [B, Lq, D]."""
    B, Lq, D = query.shape
    dh = D // n_heads
    q = F.linear(query, sd[prefix + "q.weight"])                                   # :125
    if is_cross:
        k = F.linear(key, sd[prefix + "k.weight"]); v = F.linear(value, sd[prefix + "v.weight"])   # :117-119
    else:
        k, v = F.linear(query, sd[prefix + "kv.weight"]).chunk(2, dim=-1)          # :121-123
    Lk = k.shape[1]
    heads = lambda x: x.view(B, -1, n_heads, dh).transpose(1, 2)                   # noqa: E731  :128-132
    q, k, v = heads(q), heads(k), heads(v)
    scores = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(dh))                     # :135
    if (prefix + "rel_bias.weight") in sd:
        i = torch.arange(Lq)[:, None]; j = torch.arange(Lk)[None, :]
        idx = bucket_of(j - i, num_buckets, max_distance)[None] + (torch.arange(n_heads) * num_buckets)[:, None, None]
        scores = scores + sd[prefix + "rel_bias.weight"][idx.reshape(-1), 0].view(1, n_heads, Lq, Lk)     # :137-141
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask[:, None, None, :], -1e9)      # :143-144
    if attn_mask is not None:
        scores = scores + attn_mask                                                 # :146-151
    out = torch.softmax(scores, dim=-1) @ v                                         # :153-156 (eval: dropout is the identity)
    return F.linear(out.transpose(1, 2).reshape(B, Lq, D), sd[prefix + "o.weight"])
