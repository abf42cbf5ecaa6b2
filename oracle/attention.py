"""Oracle: LLM attention over the pruned cache (test infrastructure, CPU only).

Restates the arithmetic `streaming_text_flash_attn_forward` asks of flash-attn
(`qwen2_5/language_forward.py:199-252`): GQA expansion of K/V (`repeat_kv`, `:201-202`),
softmax(Q K^T / sqrt(D)) V with a causal mask aligned to the bottom-right corner
(query row t of Tq sees keys 0 .. S-Tq+t; flash-attn >= 2.1 semantics, which HF's
`_flash_attention_forward` selects with `is_causal=True`), no dropout, no sliding
window.  Inputs are bf16, the softmax runs in fp32, the output is rounded to bf16 once.
The reference has no CPU attention (flash-attn only), so this row is the mathematical
definition; GPU parity is asserted within a bf16 tolerance, not bit-exactly.
"""
from __future__ import annotations

import math

import torch


def attention(q, k, v, causal=True, scale=None):
    """q [Hq,Tq,D], k/v [Hkv,S,D] (already rotated) -> [Tq, Hq*D] in q.dtype."""
    hq, tq, d = q.shape
    hkv, s, _ = k.shape
    g = hq // hkv
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    qf = q.float().reshape(hkv, g, tq, d)
    kf = k.float()
    vf = v.float()
    sc = torch.einsum("hgtd,hsd->hgts", qf, kf) * scale
    if causal:
        row = torch.arange(tq)[:, None] + (s - tq)
        col = torch.arange(s)[None, :]
        sc = sc.masked_fill(col > row, float("-inf"))
    p = torch.softmax(sc, dim=-1)
    o = torch.einsum("hgts,hsd->hgtd", p, vf)                  # [Hkv,G,Tq,D]
    o = o.reshape(hq, tq, d).transpose(0, 1).reshape(tq, hq * d)
    return o.to(q.dtype)


def varlen_attention(q, k, v, cu_seqlens, scale=None):
    """Non-causal packed attention (flash_attn_varlen_func): q/k/v [N,H,D]."""
    n, h, d = q.shape
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    out = torch.empty(n, h, d, dtype=torch.float32)
    cu = [int(c) for c in cu_seqlens]
    for a, b in zip(cu[:-1], cu[1:]):
        if b <= a:
            continue
        qs = q[a:b].float().transpose(0, 1)
        ks = k[a:b].float().transpose(0, 1)
        vs = v[a:b].float().transpose(0, 1)
        p = torch.softmax(qs @ ks.transpose(1, 2) * scale, dim=-1)
        out[a:b] = (p @ vs).transpose(0, 1)
    return out.to(q.dtype)
