#!/usr/bin/env python
"""Driver for an ncu launch list of ONE batched decode step (and, with PHASE=prefill, one batched ViT + chunk prefill) of the
multi-stream engine (B streams x 7B geometry, steady state S = 5789 each):
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python benchmarks/profile_multi_decode_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.multi_stream import MultiStreamEngine

B = int(os.environ.get("STREAMS", "8"))
phase = os.environ.get("PHASE", "decode")
dev = torch.device("cuda")
w = synthetic.random_weights(G, seed=0, device=dev)
multi = MultiStreamEngine(w, G, B, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G.vocab))
del w
for b, e in enumerate(multi.engines):
    e.load_state(synthetic.fast_forward(200, new_tokens=20, seed=100 + b), seed=100 + b)
base = multi.engines[0].i


def inputs(k):
    px, pr = [], []
    for b in range(B):
        p_, grid = synthetic.patchify(synthetic.frames(7919 * (b + 1) + base + k), G, dtype=torch.float32)
        px.append(p_.to(dev))
        pr.append(synthetic.round_prompt(base + k, seed=100 + b))
    return pr, px, grid


for k in range(3):
    pr, px, grid = inputs(k)
    multi.step(pr, px, grid, sync_timing=False)
torch.cuda.synchronize()
if phase == "decode":
    multi.active.fill_(1)
    multi.decode_step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    multi.decode_step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    pr, px, grid = inputs(3)
    ids_all = [e.begin_step(pr[b], grid, "")[0] for b, e in enumerate(multi.engines)]
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    multi._vit_and_prefill_all(ids_all, px, grid)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", phase)
