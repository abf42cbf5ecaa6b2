#!/usr/bin/env python
"""Developer tool: sweep the (rows per group, K-split, unroll) configurations of the decode GEMV (needs the -DSVLM_TUNE build)."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
_abi.LIB_PATH = os.path.join(ROOT, "benchmarks", "_trace", "libsvlm_tune.so")
CFG = ["4,8,2", "2,8,4", "2,8,2", "4,4,2", "4,4,4", "2,4,4", "2,2,4", "1,2,4", "1,4,4", "2,4,2", "4,2,2", "1,8,4", "2,1,4", "1,1,8"]

def timed(fn, reps=5):
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

dev = "cuda"
L = 28
for name, n, k, mode, var in [("down", 3584, 18944, 1, "SVLM_TUNE_LONGK"), ("o", 3584, 3584, 1, "SVLM_TUNE_FEWROWS"), ("qkv", 4608, 3584, 0, "SVLM_TUNE_FEWROWS")]:
    ws = [(torch.randn(n, k, device=dev) * 0.02).to(torch.bfloat16) for _ in range(L)]
    x = torch.randn(k, device=dev).to(torch.bfloat16); nw = torch.ones(k, device=dev, dtype=torch.bfloat16)
    y = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    res = {}
    for ci, cfg in enumerate(CFG):
        os.environ[var] = str(ci)
        try:
            if mode == 1:
                t = timed(lambda: [_abi.gemv_bf16(w, x, y, 1, residual=y) for w in ws]) / L
            else:
                t = timed(lambda: [_abi.gemv_bf16(w, x, y, 0, nw, 1e-6) for w in ws]) / L
            res[cfg] = round(t * 1e6, 2)
        except RuntimeError as e:
            res[cfg] = str(e)[:40]
    os.environ.pop(var)
    best = min((v, c) for c, v in res.items() if isinstance(v, float))
    print(json.dumps({"shape": name, "N": n, "K": k, "GBs_best": round(n * k * 2 / best[0] / 1e3, 1), "best": best[1], "us": res}), flush=True)
    del ws
