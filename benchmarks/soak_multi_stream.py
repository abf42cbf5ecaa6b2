#!/usr/bin/env python
"""8 streams on one GPU for `--rounds` chunks each from EMPTY caches (warm-up of the eviction policy included): per-window p50 /
p99 of the round latency, aggregate frames/s, cache lengths, CUDA-graph counts and HBM in use -- flatness check of the
multi-stream engine (BASELINE configs[3] at one GPU)."""
import argparse, json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.multi_stream import MultiStreamEngine

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=600)
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--window", type=int, default=60)
a = ap.parse_args()
dev = torch.device("cuda")
w = synthetic.random_weights(G, seed=0, device=dev)
ms = MultiStreamEngine(w, G, a.streams, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G.vocab))
del w
pool = [[synthetic.patchify(synthetic.frames(7919 * (b + 1) + k), G)[0].to(dev) for k in range(4)] for b in range(a.streams)]
grid = (1, 32, 32)
lat = []
for i in range(a.rounds):
    t0 = time.perf_counter()
    ms.step([synthetic.round_prompt(i, seed=100 + b) for b in range(a.streams)], [pool[b][i % 4] for b in range(a.streams)], grid, sync_timing=False)
    torch.cuda.current_stream().synchronize()
    lat.append(time.perf_counter() - t0)
    if (i + 1) % a.window == 0:
        wl = lat[-a.window:]
        print(json.dumps({"video_minute": (i + 1) // a.window, "rounds": i + 1, "p50_ms": round(1e3 * statistics.median(wl), 1),
                          "p99_ms": round(1e3 * sorted(wl)[int(0.99 * len(wl)) - 1], 1), "frames_per_s": round(2 * a.streams * len(wl) / sum(wl), 1),
                          "kv_len": [e.cache.seq_len for e in ms.engines], "decode_graphs": len(ms._graphs), "prefill_graphs": len(ms._pf_graphs),
                          "hbm_GB": round(torch.cuda.memory_allocated() / 1e9, 1), "hbm_reserved_GB": round(torch.cuda.memory_reserved() / 1e9, 1)}), flush=True)
