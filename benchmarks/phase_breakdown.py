#!/usr/bin/env python
"""Per-phase GPU and host time of one steady-state chunk of the 7B engine (evict / ViT / prefill / decode)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G7
from streaming_vlm_b200.engine import StreamingEngine

dev = torch.device("cuda")
w = synthetic.random_weights(G7, seed=0, device=dev)
eng = StreamingEngine(w, G7, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G7.vocab))
eng.load_state(synthetic.fast_forward(200, new_tokens=20))
rows = []
for k in range(8):
    px, grid = synthetic.patchify(synthetic.frames(eng.i), G7)
    px = px.to(dev)
    eng.phase_marks = [] if k >= 3 else None
    eng.step(synthetic.round_prompt(eng.i), px, grid, sync_timing=False)
    torch.cuda.synchronize()
    if eng.phase_marks:
        m = eng.phase_marks
        rows.append({m[i][0] + "->" + m[i + 1][0]: (m[i][1].elapsed_time(m[i + 1][1]), (m[i + 1][2] - m[i][2]) * 1e3) for i in range(len(m) - 1)})
for key in rows[0]:
    print(f"{key:32s} gpu {statistics.median(r[key][0] for r in rows):8.2f} ms   host {statistics.median(r[key][1] for r in rows):8.2f} ms")
