import sys, os, json, statistics
sys.path.insert(0, os.getcwd())
import torch
from streaming_vlm_b200 import _abi
from benchmarks.gemv_bench import timed
dev = "cuda"
for name, n, k, mode in [("o", 3584, 3584, 1), ("qkv", 4608, 3584, 0), ("gate_up", 18944, 3584, 2), ("down", 3584, 18944, 1)]:
    rows = 2 * n if mode == 2 else n
    ws = [(torch.randn(rows, k, device=dev) * 0.02).to(torch.bfloat16) for _ in range(28)]
    x = torch.randn(k, device=dev).to(torch.bfloat16)
    nw = torch.ones(k, device=dev, dtype=torch.bfloat16)
    y = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    res = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    def f(wl):
        if mode == 1:
            return lambda: [_abi.gemv_bf16(w, x, res, 1, residual=res) for w in wl]
        return lambda: [_abi.gemv_bf16(w, x, y, mode, nw, 1e-6) for w in wl]
    t_hbm = timed(f(ws)) / 28
    t_l2 = timed(f([ws[0]] * 28)) / 28
    # prefetch then gemv, alternating distinct weights: prefetch of w[i+1] issued before gemv of w[i]
    def g():
        for i, w in enumerate(ws):
            _abi.prefetch_l2(ws[(i + 1) % 28])
            if mode == 1: _abi.gemv_bf16(w, x, res, 1, residual=res)
            else: _abi.gemv_bf16(w, x, y, mode, nw, 1e-6)
    t_pf = timed(g) / 28
    print(json.dumps({"shape": name, "MB": rows * k * 2 / 1e6, "hbm_us": round(t_hbm * 1e6, 2), "same_weights_L2_us": round(t_l2 * 1e6, 2), "prefetch_next_then_gemv_us": round(t_pf * 1e6, 2)}), flush=True)
