#!/usr/bin/env python
"""Per-phase GPU and host time of one steady-state ROUND of the multi-stream engine (B streams x 7B, S ~ 5.8k each):
batched ViT / host-side eviction planning + compaction launches / batched chunk prefill / batched decode steps.
Phases are separated by synchronisations here (the product overlaps the ViT with the host planning), so the sum is an
upper bound of the round time `bench.py` reports."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G7
from streaming_vlm_b200.multi_stream import MultiStreamEngine

B = int(os.environ.get("STREAMS", "8"))
dev = torch.device("cuda")
w = synthetic.random_weights(G7, seed=0, device=dev)
multi = MultiStreamEngine(w, G7, B, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G7.vocab))
for b, e in enumerate(multi.engines):
    e.load_state(synthetic.fast_forward(200, new_tokens=20, seed=100 + b), seed=100 + b)
base = multi.engines[0].i


def inputs(k):
    px, pr = [], []
    for b in range(B):
        p_, grid = synthetic.patchify(synthetic.frames(7919 * (b + 1) + base + k), G7, dtype=torch.float32)
        px.append(p_.to(dev))
        pr.append(synthetic.round_prompt(base + k, seed=100 + b))
    return pr, px, grid


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


rows = []
for k in range(7):
    pr, px, grid = inputs(k)
    if k < 3:
        multi.step(pr, px, grid, sync_timing=False)
        continue
    torch.cuda.synchronize()
    r = {}
    a = ev(); t0 = time.perf_counter()
    video = multi.engines[0].vision_forward(torch.cat(px, dim=0), (grid[0] * B, grid[1], grid[2]))
    b_ = ev(); torch.cuda.synchronize()
    r["vit_gpu_ms"], r["vit_host_ms"] = a.elapsed_time(b_), (time.perf_counter() - t0) * 1e3
    a = ev(); t0 = time.perf_counter()
    ids_all, runs_all = [], []
    for b, e in enumerate(multi.engines):
        ids, runs = e.begin_step(pr[b], grid, "")
        ids_all.append(ids); runs_all.append(runs)
    r["evict_host_ms"] = (time.perf_counter() - t0) * 1e3
    b_ = ev(); torch.cuda.synchronize()
    r["evict_gpu_ms"] = a.elapsed_time(b_)
    a = ev(); t0 = time.perf_counter()
    multi._vit_and_prefill_all(ids_all, px, grid, video)
    b_ = ev(); torch.cuda.synchronize()
    r["prefill_gpu_ms"], r["prefill_host_ms"] = a.elapsed_time(b_), (time.perf_counter() - t0) * 1e3
    multi.active.fill_(1)
    seqs = [list(i) for i in ids_all]
    dec = []
    for step in range(19):
        a = ev()
        multi.decode_step()
        b_ = ev(); torch.cuda.synchronize()
        dec.append(a.elapsed_time(b_))
        for e in multi.engines:
            e.cache.seq_len += 1
            e.cache._seen_tokens += 1
    multi._tok_pinned.copy_(multi.tok)
    torch.cuda.synchronize()
    r["decode_step_ms"] = statistics.median(dec)
    r["decode_19_ms"] = sum(dec)
    for b, e in enumerate(multi.engines):
        seqs[b] += [1000 + j for j in range(19)] + [int(multi._tok_pinned[b])]
        e.finish_step(ids_all[b], seqs[b], runs_all[b])
    rows.append(r)
out = {key: round(statistics.median(r[key] for r in rows), 3) for key in rows[0]}
out["streams"] = B
out["sum_gpu_ms"] = round(out["vit_gpu_ms"] + out["evict_gpu_ms"] + out["prefill_gpu_ms"] + out["decode_19_ms"], 2)
print(json.dumps(out))
