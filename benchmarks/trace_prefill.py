#!/usr/bin/env python
"""Developer tool: per-tile clock64 timeline of the prefill attention kernel (-DSVLM_TRACE build)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
_abi.LIB_PATH = os.path.join(ROOT, "benchmarks", "_trace", os.environ.get("SVLM_TRACE_LIB", "libsvlm_trace.so"))
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
tq, splits = 276, 2
dev = torch.device("cuda")
c = StreamingCache(G, n + 512, dev)
c.arena[0].normal_(0, 0.5)
HQ, D = G.num_heads, G.head_dim
mt = -(-tq // 18)
trace = torch.zeros(splits * 4 * mt * 32 * 16 + 4096, dtype=torch.int64, device=dev)
os.environ["SVLM_TRACE_PTR"] = str(trace.data_ptr())
q = torch.randn(tq, HQ, D, device=dev).to(torch.bfloat16); o = torch.empty_like(q)
ws = torch.zeros(_abi.attn_workspace_bytes(tq, HQ, D, splits) // 4 + 16, dtype=torch.float32, device=dev)
for it in range(3):
    _abi.attn_causal(c.view, 0, it, q, o, tq, HQ, n, 1 / math.sqrt(D), splits, ws)
torch.cuda.synchronize()
t = trace.view(-1, 32, 16).cpu()
cnt = (t[:, :, 1] > 0).sum(1)
cta = int(cnt.argmax())
tt = t[cta]
nt = int(cnt[cta])
base = int(tt[6, 0])
print("cta", cta, "tiles", nt)
for j in range(6, min(nt, 9)):
    print(j, " ".join(f"w{w + 2}:{int(tt[j, 2 * w]) - base}->{int(tt[j, 2 * w + 1]) - base}({int(tt[j, 2 * w + 1]) - int(tt[j, 2 * w])})" for w in range(8)))
print("cycles/tile:", (int(tt[nt - 1, 1]) - int(tt[2, 1])) / (nt - 3))
