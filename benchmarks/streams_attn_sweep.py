#!/usr/bin/env python
"""Developer sweep: split count of the batched (multi-stream) decode attention, 7B geometry, B streams at kv_len n."""
import json, math, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5789
for B in [int(a) for a in sys.argv[2:]] or (2, 4, 8, 16):
    caches = [StreamingCache(G, 8192, dev) for _ in range(B)]
    for c in caches:
        c.arena[0].normal_(0, 0.5)
    kv = torch.full((B,), n, dtype=torch.int32, device=dev)
    one = list(kv.split(1))
    slots = _abi.make_slots(caches, one, one, one)
    q = torch.randn(B, G.num_heads, G.head_dim, device=dev).to(torch.bfloat16)
    o = torch.empty_like(q)
    byt = B * n * 2 * G.num_kv_heads * G.head_dim * 2
    for splits in (2, 3, 4, 5, 6, 8, 9, 10, 12, 16, 23, 37):
        ws = torch.zeros(_abi.attn_workspace_bytes(B, G.num_heads, G.head_dim, splits) // 4 + 16, dtype=torch.float32, device=dev)
        def f():
            for li in range(G.num_layers):
                _abi.attn_decode_streams(caches[0].view, slots, li, q, o, G.num_heads, 8192, 1 / math.sqrt(128), splits, ws)
        f(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            f()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3 / G.num_layers)
        t = statistics.median(ts)
        print(json.dumps({"streams": B, "kv_len": n, "splits": splits, "ctas": splits * G.num_kv_heads * B, "us": round(t * 1e6, 2),
                          "GBs": round(byt / t / 1e9, 1), "frac_hbm": round(byt / t / 1e9 / 6533.8, 3)}), flush=True)
    del caches
