#!/usr/bin/env python
"""Developer probe: T = 1..16 tokens through the decoder-layer projections, 28 distinct weight sets (>> L2), CUDA-graph replay.
cuBLAS (F.linear, epilogue ops NOT included) vs svlm_gemm_tokens (bias / residual / SwiGLU fused): how close is a B-stream batched
decode step to the HBM floor of ONE weight sweep?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from streaming_vlm_b200 import _abi
from benchmarks.gemv_bench import timed

dev = "cuda"
PEAK = 6533.8e9
ms = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]
for name, n, k, epi in [("qkv", 4608, 3584, _abi.EPI_NONE), ("o", 3584, 3584, _abi.EPI_RESIDUAL), ("gate_up", 37888, 3584, _abi.EPI_SWIGLU),
                        ("down", 3584, 18944, _abi.EPI_RESIDUAL), ("lm_head", 152064, 3584, _abi.EPI_NONE)]:
    copies = 28 if n < 100000 else 4
    ws = [(torch.randn(n, k, device=dev) * 0.02).to(torch.bfloat16) for _ in range(copies)]
    byt = n * k * 2
    for m in ms:
        x = torch.randn(m, k, device=dev).to(torch.bfloat16)
        y = torch.empty(m, n // 2 if epi == _abi.EPI_SWIGLU else n, device=dev, dtype=torch.bfloat16)
        res = torch.randn_like(y) if epi == _abi.EPI_RESIDUAL else None
        bias = torch.randn(n, device=dev).to(torch.bfloat16) if name == "qkv" else None
        tb = timed(lambda: [F.linear(x, w) for w in ws]) / copies
        row = {"shape": name, "T": m, "cublas_us": round(tb * 1e6, 2), "cublas_frac_hbm": round(byt / tb / PEAK, 3)}
        for s in ([0] if n > 8000 else [0, 3, 4, 5, 6, 7, 8]):
            to = timed(lambda: [_abi.gemm_tokens(x, w, y, bias=bias, residual=res, epilogue=epi, splits=s) for w in ws]) / copies
            row[f"own_s{s}_us"] = round(to * 1e6, 2)
            if s == 0:
                row["own_frac_hbm"] = round(byt / to / PEAK, 3)
        if epi != _abi.EPI_RESIDUAL:
            nw = torch.ones(k, device=dev, dtype=torch.bfloat16)
            to = timed(lambda: [_abi.gemm_tokens(x, w, y, bias=bias, residual=res, epilogue=epi, norm_weight=nw) for w in ws]) / copies
            row["own_fused_norm_us"] = round(to * 1e6, 2)
            st = torch.rand(k // 128, m, device=dev) * 128
            to = timed(lambda: [_abi.gemm_tokens(x, w, y, bias=bias, residual=res, epilogue=epi, norm_weight=nw, stats_in=st) for w in ws]) / copies
            row["own_fused_norm_stats_in_us"] = round(to * 1e6, 2)
        print(json.dumps(row), flush=True)
    del ws
