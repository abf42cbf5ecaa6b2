#!/usr/bin/env python
"""BASELINE config 2: one 7B-architecture random-init stream, synthetic 2 fps video, per-second commentary decode, run for
`--chunks` one-second chunks (3600 = one hour of video) from an EMPTY cache through the whole warm-up of the
sink/text-window/vision-window policy into steady state.  Reports per-minute p50 / p99 chunk latency, frames/s and decode
tokens/s to show that latency stays flat once the cache has filled, plus the retained cache length."""
import argparse, json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G7
from streaming_vlm_b200.engine import StreamingEngine

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=3600)
ap.add_argument("--window", type=int, default=60)
args = ap.parse_args()
dev = torch.device("cuda")
w = synthetic.random_weights(G7, seed=0, device=dev)
eng = StreamingEngine(w, G7, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G7.vocab))
del w
pool = [synthetic.patchify(synthetic.frames(k), G7)[0].pin_memory() for k in range(16)]     # 16 distinct synthetic chunks, recycled
grid = (1, 32, 32)
lat, toks, kv = [], [], []
t_all = time.perf_counter()
for i in range(args.chunks):
    t0 = time.perf_counter()
    px = pool[i % 16].to(dev, non_blocking=True)
    r = eng.step(synthetic.round_prompt(i), px, grid, sync_timing=False)
    torch.cuda.current_stream().synchronize()
    lat.append(time.perf_counter() - t0)
    toks.append(min(len(r["new_tokens"]), 20))
    kv.append(r["cache_len"])
    if (i + 1) % args.window == 0:
        wl = lat[-args.window:]
        print(json.dumps({"video_minute": (i + 1) // args.window, "chunks": i + 1, "p50_ms": round(1e3 * statistics.median(wl), 2),
                          "p99_ms": round(1e3 * sorted(wl)[int(0.99 * len(wl)) - 1], 2), "max_ms": round(1e3 * max(wl), 2),
                          "frames_per_s": round(2 * len(wl) / sum(wl), 2), "decode_tok_s": round(sum(toks[-args.window:]) / sum(wl), 1),
                          "kv_len": kv[-1], "compactions": eng.cache.compactions, "graphs": len(eng._pf_graphs) + len(eng._graphs)}), flush=True)
steady = lat[len(lat) // 2:]
print(json.dumps({"summary": True, "chunks": args.chunks, "wall_s": round(time.perf_counter() - t_all, 1),
                  "steady_p50_ms": round(1e3 * statistics.median(steady), 2), "steady_p99_ms": round(1e3 * sorted(steady)[int(0.99 * len(steady)) - 1], 2),
                  "first_half_p50_ms": round(1e3 * statistics.median(lat[:len(lat) // 2]), 2), "kv_len_max": max(kv), "kv_len_final": kv[-1],
                  "realtime_factor": round(args.chunks / (time.perf_counter() - t_all), 2)}), flush=True)
