#!/usr/bin/env python
"""Developer probe: M = 1..16 rows through the decoder-layer projections, 28 distinct weight sets (>> L2), CUDA-graph replay.
cuBLAS (F.linear) vs svlm_gemm_bf16: how close is a B-stream batched decode step to the HBM floor of ONE weight sweep?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from streaming_vlm_b200 import _abi
from benchmarks.gemv_bench import timed

dev = "cuda"
for name, n, k in [("qkv", 4608, 3584), ("o", 3584, 3584), ("gate_up", 37888, 3584), ("down", 3584, 18944)]:
    ws = [(torch.randn(n, k, device=dev) * 0.02).to(torch.bfloat16) for _ in range(28)]
    byt = n * k * 2
    for m in (1, 2, 4, 8, 16):
        x = torch.randn(m, k, device=dev).to(torch.bfloat16)
        y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        tb = timed(lambda: [F.linear(x, w) for w in ws]) / 28
        to = timed(lambda: [_abi.gemm_bf16(x, w, None, y) for w in ws]) / 28
        print(json.dumps({"shape": name, "M": m, "cublas_us": round(tb * 1e6, 2), "cublas_frac_hbm": round(byt / tb / 6533.8e9, 3),
                          "svlm_gemm_us": round(to * 1e6, 2), "svlm_frac_hbm": round(byt / to / 6533.8e9, 3)}), flush=True)
    del ws
