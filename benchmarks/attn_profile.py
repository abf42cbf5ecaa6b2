#!/usr/bin/env python
"""Small driver for ncu captures (`ncu --set full -k regex:... python benchmarks/attn_profile.py`): a few launches of every kernel
family at the steady-state shapes of the 7B geometry (S = 5789): decode / prefill attention, ViT windowed and full attention + RoPE,
compaction, the decode GEMVs incl. the sampling lm_head, the embedding splice and the patch-merger GEMM (CTA-pair kernel)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.engine import decode_split_count
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
splits = decode_split_count(n, G.num_kv_heads)
ws = torch.zeros(max(_abi.attn_workspace_bytes(300, HQ, D, 4), _abi.attn_workspace_bytes(1, HQ, D, splits)) // 4 + 16, dtype=torch.float32, device=dev)
kv = torch.tensor([n], dtype=torch.int32, device=dev)
qp = torch.randn(300, HQ, D, device=dev).to(torch.bfloat16); op = torch.empty_like(qp)
w = (torch.randn(2 * G.mlp, G.hidden, device=dev) * 0.02).to(torch.bfloat16)
x = torch.randn(G.hidden, device=dev).to(torch.bfloat16); nw = torch.ones_like(x)
act = torch.empty(G.mlp, dtype=torch.bfloat16, device=dev)
# ViT: one chunk = 1024 patch tokens, 16 heads of 80; 16 windows of 64 (windowed layers) or one sequence of 1024 (full layers)
vh, vd = G.vit_heads, G.vit_hidden // G.vit_heads
qkv = torch.randn(1024, 3, vh, vd, device=dev).to(torch.bfloat16)
ov = torch.empty(1024, vh, vd, dtype=torch.bfloat16, device=dev)
cu_w = torch.arange(0, 1025, 64, dtype=torch.int32, device=dev)
cu_f = torch.tensor([0, 1024], dtype=torch.int32, device=dev)
cosv = torch.randn(1024, vd, device=dev); sinv = torch.randn(1024, vd, device=dev)
# patch merger fc1 (256 x 5120 x 5120, GELU) on the persistent GEMM; lm_head with the reference's sampling config; embedding splice
ma = torch.randn(256, 5120, device=dev).to(torch.bfloat16)
mw = (torch.randn(5120, 5120, device=dev) * 0.01).to(torch.bfloat16)
mb = torch.zeros(5120, dtype=torch.bfloat16, device=dev)
mo = torch.empty(256, 5120, dtype=torch.bfloat16, device=dev)
lm = (torch.randn(G.vocab, G.hidden, device=dev) * 0.02).to(torch.bfloat16)
st = dict(key=torch.zeros(1, dtype=torch.int64, device=dev), tok=torch.zeros(1, dtype=torch.int64, device=dev),
          cnt=torch.zeros(3, dtype=torch.int32, device=dev), rng=torch.zeros(4, dtype=torch.int32, device=dev),
          seen=torch.zeros(G.vocab, dtype=torch.uint8, device=dev))
ids = torch.randint(0, 150000, (304,), device=dev)
ids[20:276] = 151656
vid = torch.randn(256, G.hidden, device=dev).to(torch.bfloat16)
xe = torch.empty(304, G.hidden, dtype=torch.bfloat16, device=dev)
for it in range(3):
    for li in range(4):
        _abi.attn_causal(c.view, 0, li, q1, o1, 1, HQ, c.capacity, scale, splits, ws, kv_len_dev=kv)
    for li in range(4):
        _abi.attn_causal(c.view, 0, li, qp, op, 300, HQ, n, scale, 2, ws)
    _abi.kv_compact_rope(c.view, 0, [(0, 0, n)], c.pos, c.cos_table, c.sin_table, c.section)
    _abi.gemv_bf16(w, x, act, 2, nw, 1e-6)
    _abi.vit_rope(qkv, cosv, sinv, 1024, vh, vd)
    _abi.vit_attn(qkv, cu_w, ov, 1024, vh, vd, vd ** -0.5)
    _abi.vit_attn(qkv, cu_f, ov, 1024, vh, vd, vd ** -0.5)
    _abi.gemm_ex(ma, mw, mo, bias=mb, epilogue=_abi.EPI_GELU)
    _abi.lm_head_sample(lm, x, nw, 1e-6, None, st["key"], st["tok"], st["cnt"][0:1], st["cnt"][1:2], st["cnt"][2:3], seen=st["seen"],
                        repetition_penalty=1.05, temperature=0.9, do_sample=True, rng_state=st["rng"])
    _abi.embed_tokens(ids, 304, lm, vid, 256, xe, None, 151656)
torch.cuda.synchronize()
print("done", n, splits)
