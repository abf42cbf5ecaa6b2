#!/usr/bin/env python
"""Developer sweep: single-stream split-KV decode attention, split count x merge path.  Split counts 2 / 4 / 8 / 16 merge inside ONE
thread-block cluster (push over distributed shared memory, no merge launch); every other count uses the separate merge kernel.
SVLM_DECODE_CLUSTER=1 forces the merge kernel for all counts (run the script twice for the A/B)."""
import json, math, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.engine import decode_split_count
from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache

dev = torch.device("cuda")
HQ, D, L = G.num_heads, G.head_dim, G.num_layers
lens = [int(a) for a in sys.argv[1:]] or [4096, 5789, 8192, 16384, 32768, 65536]
for n in lens:
    c = StreamingCache(G, ((n + 640) // 128) * 128, dev)
    c.arena[0].normal_(0, 0.5)
    c.seq_len = n
    kv = torch.tensor([n], dtype=torch.int32, device=dev)
    q = torch.randn(1, HQ, D, device=dev).to(torch.bfloat16)
    o = torch.empty_like(q)
    byt = n * 2 * G.num_kv_heads * D * 2
    heur = decode_split_count(n, G.num_kv_heads, 296 // G.num_kv_heads)
    row = {"kv_len": n, "merge": "kernel only" if os.environ.get("SVLM_DECODE_CLUSTER") == "1" else "cluster for 2/4/8/16", "heuristic_splits": heur}
    for splits in sorted({4, 8, 16, 23, 32, 37, heur}):
        if splits > (n + 127) // 128:
            continue
        ws = torch.zeros(_abi.attn_workspace_bytes(1, HQ, D, splits) // 4 + 16, dtype=torch.float32, device=dev)

        def f():
            for li in range(L):
                _abi.attn_causal(c.view, 0, li, q, o, 1, HQ, c.capacity, 1 / math.sqrt(D), splits, ws, kv_len_dev=kv)
        f(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            f()
        gr.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3 / L)
        t = statistics.median(ts)
        row[f"s{splits}_us"] = round(t * 1e6, 2)
        row[f"s{splits}_frac"] = round(byt / t / 1e9 / 6533.8, 3)
    print(json.dumps(row), flush=True)
    del c
