#!/usr/bin/env python
"""Decode GEMV microbench: each projection shape of the 7B decoder layer, 28 distinct weight sets (>> L2), CUDA-graph
replay timed with events; svlm_gemv_bf16 vs cuBLAS (torch.nn.functional.linear, M = 1)."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from streaming_vlm_b200 import _abi


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    gr.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)


def main():
    _abi.build()
    dev = "cuda"
    L = 28
    for name, n, k, mode in [("qkv", 4608, 3584, 0), ("o", 3584, 3584, 1), ("gate_up", 18944, 3584, 2), ("down", 3584, 18944, 1),
                             ("lm_head", 152064, 3584, 3)]:
        reps = L if mode != 3 else 4
        rows = 2 * n if mode == 2 else n
        ws = [(torch.randn(rows, k, device=dev) * 0.02).to(torch.bfloat16) for _ in range(reps)]
        x = torch.randn(k, device=dev).to(torch.bfloat16)
        nw = torch.ones(k, device=dev, dtype=torch.bfloat16)
        y = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        res = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        key = torch.zeros(1, dtype=torch.int64, device=dev); tok = torch.zeros(1, dtype=torch.int64, device=dev)
        cnt = torch.zeros(3, dtype=torch.int32, device=dev)
        if mode == 3:
            f = lambda: [_abi.lm_head_argmax(w, x, nw, 1e-6, None, key, tok, cnt[0:1], cnt[1:2], cnt[2:3]) for w in ws]
        elif mode == 1:
            f = lambda: [_abi.gemv_bf16(w, x, res, 1, residual=res) for w in ws]
        else:
            f = lambda: [_abi.gemv_bf16(w, x, y, mode, nw, 1e-6) for w in ws]
        t = timed(f) / reps
        x2 = x[None]
        tb = timed(lambda: [F.linear(x2, w) for w in ws]) / reps
        byt = rows * k * 2
        print(json.dumps({"shape": name, "rows": rows, "K": k, "svlm_us": round(t * 1e6, 2), "svlm_GBs": round(byt / t / 1e9, 1),
                          "cublas_us": round(tb * 1e6, 2), "cublas_GBs": round(byt / tb / 1e9, 1)}), flush=True)
        del ws


if __name__ == "__main__":
    main()
