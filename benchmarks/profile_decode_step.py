#!/usr/bin/env python
"""Driver for an ncu launch list of ONE decode forward of the final engine (7B geometry, steady state S = 5789):
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python benchmarks/profile_decode_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.engine import StreamingEngine

dev = torch.device("cuda")
w = synthetic.random_weights(G, seed=0, device=dev)
eng = StreamingEngine(w, G, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G.vocab))
del w
eng.load_state(synthetic.fast_forward(200, new_tokens=20, seed=0), seed=0)
for k in range(3):
    fr = synthetic.frames(eng.i).permute(0, 2, 3, 1).contiguous().to(dev)
    eng.step(synthetic.round_prompt(eng.i), None, None, sync_timing=False, frames=fr)
for _ in range(2):
    eng.decode_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.decode_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done, kv_len", eng.cache.seq_len)
