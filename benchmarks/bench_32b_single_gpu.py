#!/usr/bin/env python
"""BASELINE configs[4] data point at TP = 1: the Qwen2.5-VL-32B-architecture (64 layers, hidden 5120, 40 q / 8 kv heads,
65 GB of bf16 weights) fits ONE B200 (180 GB), so the long-stream steady state can be measured without tensor parallelism --
same engine, same kernels (G = 5 query heads per kv head), same synthetic stream as bench.py.  The head-sharded TP mode
(NCCL all-reduce after o_proj and down_proj) is not built; this is the number it would have to beat per GPU."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_32B as G
from streaming_vlm_b200.engine import StreamingEngine

steps, warmup = 6, 3
dev = torch.device("cuda")
w = synthetic.random_weights(G, seed=0, device=dev)
eng = StreamingEngine(w, G, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G.vocab))
del w
eng.load_state(synthetic.fast_forward(200, new_tokens=20, seed=0), seed=0)
base = eng.i
frames = [synthetic.frames(base + k).permute(0, 2, 3, 1).contiguous().to(dev) for k in range(steps + warmup)]
prompts = [synthetic.round_prompt(base + k) for k in range(steps + warmup)]
lat, toks = [], 0
for k in range(warmup):
    eng.step(prompts[k], None, None, sync_timing=False, frames=frames[k])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(warmup, warmup + steps):
    t0 = time.perf_counter()
    r = eng.step(prompts[k], None, None, sync_timing=False, frames=frames[k])
    torch.cuda.current_stream().synchronize()
    lat.append(time.perf_counter() - t0)
    toks += min(len(r["new_tokens"]), 20)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3
wbytes = sum(v.numel() * 2 for L in eng.layers for v in L.values()) + eng.lm_head.numel() * 2
print(json.dumps({"geometry": G.name, "tp": 1, "frames_per_s": 2 * steps / t, "ms_per_chunk": 1e3 * t / steps, "decode_tok_s": toks / t,
                  "p50_chunk_ms": 1e3 * statistics.median(lat), "kv_len": eng.cache.seq_len, "weights_GB_swept_per_token": wbytes / 1e9,
                  "hbm_floor_ms_per_token": wbytes / 6533.8e9 * 1e3, "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
