#!/usr/bin/env python
"""svlm_gemm_bf16 (tcgen05, one 128xBN tile per CTA) vs cuBLAS on a few shapes."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
_abi.build()
dev = "cuda"
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)
for m, n, k in [(8192, 8192, 8192), (4096, 4096, 4096), (256, 5120, 5120), (1024, 3840, 1280), (278, 37888, 3584)]:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16); w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    c = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    t1 = timed(lambda: _abi.gemm_bf16(a, w, None, c))
    t2 = timed(lambda: torch.nn.functional.linear(a, w))
    fl = 2.0 * m * n * k
    print(json.dumps({"M": m, "N": n, "K": k, "svlm_us": round(t1 * 1e6, 1), "svlm_TFs": round(fl / t1 / 1e12, 1), "cublas_us": round(t2 * 1e6, 1), "cublas_TFs": round(fl / t2 / 1e12, 1)}), flush=True)
