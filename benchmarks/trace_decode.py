#!/usr/bin/env python
"""Developer tool: per-tile clock64 timeline of the decode attention kernel (needs the -DSVLM_TRACE build of the
library in benchmarks/_trace/libsvlm_trace.so; the product library carries no trace code)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import _abi
_abi.LIB_PATH = os.path.join(ROOT, "benchmarks", "_trace", "libsvlm_trace.so")
from streaming_vlm_b200.config import QWEN25VL_7B as G
from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
c = StreamingCache(G, n + 512, dev)
c.arena[0].normal_(0, 0.5)
HQ, D = G.num_heads, G.head_dim
splits = 37
trace = torch.zeros(splits * 4 * 32 * 8, dtype=torch.int64, device=dev)
os.environ["SVLM_TRACE_PTR"] = str(trace.data_ptr())
q1 = torch.randn(1, HQ, D, device=dev).to(torch.bfloat16); o1 = torch.empty_like(q1)
ws = torch.zeros(_abi.attn_workspace_bytes(1, HQ, D, splits) // 4 + 16, dtype=torch.float32, device=dev)
kv = torch.tensor([n], dtype=torch.int32, device=dev)
for it in range(3):
    _abi.attn_causal(c.view, 0, it, q1, o1, 1, HQ, c.capacity, 1 / math.sqrt(D), splits, ws, kv_len_dev=kv)
torch.cuda.synchronize()
t = trace.view(4, splits, 32, 8).cpu()
names = ["sm:loop_top", "sm:s_full_seen", "sm:tmem_ld_done", "sm:bar_done", "sm:p_arrived", "sm:exp_done", "sm:rescale_done", "sm:stores_done"]
for cta in [(0, 0)]:
    tt = t[cta[0], cta[1]]
    ntile = int((tt[:, 1] > 0).sum())
    base = int(tt[0, 0])
    print(f"CTA head={cta[0]} split={cta[1]} tiles={ntile}  (cycles since first TMA issue)")
    for j in range(min(ntile, 14)):
        print(j, " ".join(f"{names[e].split(':')[1]}={int(tt[j, e]) - base:6d}" for e in [0, 1, 2, 3, 5, 6, 7, 4]))
    if ntile > 3:
        per = (int(tt[ntile - 1, 4]) - int(tt[1, 4])) / (ntile - 2)
        print("steady-state cycles per tile:", per)
