#!/usr/bin/env python
"""Small driver for ncu captures: a few launches of the decode / prefill attention, the compaction and a decode GEMV
at the steady-state shape of the 7B geometry (S = 5789)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache

_abi.build()
dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5789
c = StreamingCache(G, ((n + 640) // 128) * 128, dev)
c.arena[0].normal_(0, 0.5)
c.seq_len = n
HQ, D, L = G.num_heads, G.head_dim, G.num_layers
scale = 1 / math.sqrt(D)
q1 = torch.randn(1, HQ, D, device=dev).to(torch.bfloat16); o1 = torch.empty_like(q1)
splits = max(1, min((c.capacity + 127) // 128, 148 // G.num_kv_heads))
ws = torch.zeros(_abi.attn_workspace_bytes(300, HQ, D, 4) // 4 + 16, dtype=torch.float32, device=dev)
kv = torch.tensor([n], dtype=torch.int32, device=dev)
qp = torch.randn(300, HQ, D, device=dev).to(torch.bfloat16); op = torch.empty_like(qp)
w = (torch.randn(2 * G.mlp, G.hidden, device=dev) * 0.02).to(torch.bfloat16)
x = torch.randn(G.hidden, device=dev).to(torch.bfloat16); nw = torch.ones_like(x)
act = torch.empty(G.mlp, dtype=torch.bfloat16, device=dev)
for it in range(3):
    for li in range(4):
        _abi.attn_causal(c.view, 0, li, q1, o1, 1, HQ, c.capacity, scale, splits, ws, kv_len_dev=kv)
    for li in range(4):
        _abi.attn_causal(c.view, 0, li, qp, op, 300, HQ, n, scale, 2, ws)
    _abi.kv_compact_rope(c.view, 0, [(0, 0, n)], c.pos, c.cos_table, c.sin_table, c.section)
    _abi.gemv_bf16(w, x, act, 2, nw, 1e-6)
torch.cuda.synchronize()
print("done", n, splits)
