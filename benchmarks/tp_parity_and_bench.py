#!/usr/bin/env python
"""Head-sharded tensor parallelism (BASELINE configs[4]): row-parallel projections reduced by our own one-shot all-reduce kernels over
NVLink symmetric memory (svlm_tp_gemv_allreduce / svlm_tp_allreduce_rows), CUDA graphs on; NCCL only sets the group up.  Launch with
torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/tp_parity_and_bench.py parity
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/tp_parity_and_bench.py bench32b

parity  : tiny geometry; the TP engine of every rank is teacher-forced with the tokens of an unsharded engine on rank 0 and must
          reproduce its stream (ids, cache lengths, greedy ids up to near-ties; every rank checks its own copy).
bench32b: Qwen2.5-VL-32B-architecture steady state, weights sharded over the ranks; prints frames/s of the ONE stream.
"""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from streaming_vlm_b200 import parallel, synthetic
from streaming_vlm_b200.config import QWEN25VL_32B, TINY
from streaming_vlm_b200.engine import StreamingEngine

mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
use_graph = os.environ.get("SVLM_TP_EAGER", "0") != "1"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
warm = torch.ones(1, device=dev)
dist.all_reduce(warm)                                   # communicator set-up outside any CUDA-graph capture

if mode == "parity":
    geo = TINY
    w = synthetic.random_weights(geo, seed=0)
    sup = synthetic.suppress_mask(geo.vocab)
    kw = dict(capacity=4096, window_size=4, text_round=4, text_sink=16, text_sliding_window=24, suppress=sup, device=dev)
    mine, local_geo = parallel.shard_llm_weights(w, geo, rank, world)
    tp = StreamingEngine({k: v.clone() for k, v in mine.items()}, local_geo, tp_group=dist.group.WORLD, use_graph=use_graph, **kw)
    full = StreamingEngine({k: v.clone() for k, v in w.items()}, geo, **kw) if rank == 0 else None
    mism = total = 0
    for i in range(12):
        if rank == 0:
            print("chunk", i, flush=True)
        px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
        prompt = synthetic.round_prompt(i, 16, 6)
        box = [None]
        if rank == 0:
            r_full = full.step(prompt, px.to(dev), grid)
            box = [(r_full["new_tokens"], r_full["ids"], r_full["cache_len"])]
        dist.broadcast_object_list(box, src=0)
        want_new, want_ids, want_len = box[0]
        r = tp.step(prompt, px.to(dev), grid, forced=want_new)
        assert r["ids"] == want_ids and r["cache_len"] == want_len, f"rank {rank} chunk {i}"
        for a, b in zip(tp.last_argmax, want_new):
            total += 1
            mism += a != b
    assert tp.cache.compactions >= 6
    assert mism <= max(2, total // 20), (mism, total)
    # every rank must hold the same stream: the all-reduce sums in rank order on every rank, so x is bit-identical everywhere
    digest = [None] * world
    dist.all_gather_object(digest, (tp.prev_ids.tolist(), tp.cache.seq_len))
    assert all(d == digest[0] for d in digest), "ranks disagree on the generated stream"
    print(json.dumps({"mode": "parity", "rank": rank, "tp": world, "chunks": 12, "tokens": total, "argmax_mismatches_vs_unsharded": mism,
                      "compactions": tp.cache.compactions, "kv_heads_local": local_geo.num_kv_heads, "all_reduces": tp.tp.all_reduces,
                      "cuda_graphs": use_graph}), flush=True)
elif mode == "kernel":
    # the fused row-parallel GEMV + one-shot all-reduce + residual kernel on its own, 32B shapes of this TP degree: 64 layers x
    # (o_proj, down_proj) launches per CUDA-graph replay over distinct weights (>> L2), CUDA events, max over ranks
    from streaming_vlm_b200 import _abi
    geo = QWEN25VL_32B
    ctx = parallel.TpContext(dist.group.WORLD, geo.hidden, max_rows=1024, device=dev)
    hq, mlp = geo.num_heads * geo.head_dim // world, geo.mlp // world
    n_l = 24
    wo = [(torch.randn(geo.hidden, hq, device=dev) * 0.02).to(torch.bfloat16) for _ in range(n_l)]
    wd = [(torch.randn(geo.hidden, mlp, device=dev) * 0.02).to(torch.bfloat16) for _ in range(n_l)]
    xo, xd = torch.randn(hq, device=dev).to(torch.bfloat16), torch.randn(mlp, device=dev).to(torch.bfloat16)
    x = torch.randn(geo.hidden, device=dev).to(torch.bfloat16)
    res = {}
    for name, which, ws_, xin in (("o_proj", 0, wo, xo), ("down_proj", 1, wd, xd)):
        def body():
            for li in range(n_l):
                _abi.tp_gemv_allreduce(ctx.ref, which, ws_[li], xin, x, x)
        body(); torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            body()
        gr.replay(); torch.cuda.synchronize(); dist.barrier()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3 / n_l)
        t = torch.tensor([statistics.median(ts)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wbytes = ws_[0].numel() * 2
        res[name] = {"us_per_launch": float(t) * 1e6, "weight_bytes_per_rank": wbytes, "hbm_floor_us": wbytes / 6533.8e9 * 1e6,
                     "frac_of_hbm_floor": wbytes / 6533.8e9 / float(t), "nvlink_bytes_pulled": ctx.nvlink_bytes_per_decode_allreduce(),
                     "nvlink_floor_us_at_770GBs": ctx.nvlink_bytes_per_decode_allreduce() / 770e9 * 1e6}
    if rank == 0:
        print(json.dumps({"mode": "kernel", "geometry": geo.name, "tp": world, "svlm_tp_gemv_allreduce": res}), flush=True)
else:
    geo = QWEN25VL_32B
    w = synthetic.random_weights(geo, seed=0, device=dev)
    mine, local_geo = parallel.shard_llm_weights(w, geo, rank, world)
    del w
    torch.cuda.empty_cache()
    eng = StreamingEngine(mine, local_geo, capacity=8192, device=dev, suppress=synthetic.suppress_mask(geo.vocab), tp_group=dist.group.WORLD, use_graph=use_graph)
    del mine
    eng.load_state(synthetic.fast_forward(200, new_tokens=20, seed=0), seed=0)
    steps, warmup = 6, 3
    base = eng.i
    frames = [synthetic.frames(base + k).permute(0, 2, 3, 1).contiguous().to(dev) for k in range(steps + warmup)]
    prompts = [synthetic.round_prompt(base + k) for k in range(steps + warmup)]
    for k in range(warmup):
        eng.step(prompts[k], None, None, sync_timing=False, frames=frames[k])
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    toks = 0
    for k in range(warmup, warmup + steps):
        r = eng.step(prompts[k], None, None, sync_timing=False, frames=frames[k])
        toks += min(len(r["new_tokens"]), 20)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ar_per_token = 2 * geo.num_layers
        print(json.dumps({"mode": "bench32b", "geometry": geo.name, "tp": world, "frames_per_s": 2 * steps / float(t), "ms_per_chunk": 1e3 * float(t) / steps,
                          "decode_tok_s": toks / float(t), "kv_len": eng.cache.seq_len, "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                          "cuda_graphs": use_graph, "all_reduces_per_token": ar_per_token,
                          "nvlink_bytes_pulled_per_token_per_rank": ar_per_token * eng.tp.nvlink_bytes_per_decode_allreduce()}), flush=True)
dist.barrier()
dist.destroy_process_group()
