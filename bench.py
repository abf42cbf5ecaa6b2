#!/usr/bin/env python
"""Benchmark of the StreamingVLM hot path on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one steady-state 1-second chunk of ONE video stream through the whole path: eviction
(sink / text-window / vision-window compaction + contiguous-RoPE re-rotation), ViT on 2 new 448x448
frames, chunk prefill of the ~300 uncached tokens against the ~5.5k-token pruned cache, and 20 greedily
decoded commentary tokens, on a random-init Qwen2.5-VL-7B-architecture model in bf16.  The session is put
into the steady state of a long stream (round 200 of the default 16/16/512/512 policy) by replaying the
host-side bookkeeping and filling the arena; nothing inside the timed region is skipped.

The timed call is the REFERENCE-FACING one: a stock transformers Qwen2_5_VLForConditionalGeneration (7B architecture, random
init) patched with `convert_qwen2_5_to_streaming`, driven by the reference's own loop body (`inference/inference.py:319,
397-416, 440-463`): `process_past_kv(...)` -> `model.generate(**inputs, past_key_values=..., streaming_args=...,
do_sample=True, temperature=0.9, repetition_penalty=1.05, max_new_tokens=20)`.  `value`: frames resident in HBM; `e2e`: uint8
frames in pinned host memory, ids / tokens through the host as the loop does.  `engine_api`: the same chunks through
`StreamingEngine.step` directly (greedy), for comparison.

N > 1 (torchrun): every rank runs its own independent stream (no data-path collective; the reference's
one-process-per-GPU sharding, eval/livesports3kcc/distributed_generate_streaming.py:53-62) -> weak scaling.

`--impl reference`: the reference's path restated on the CPU (the oracle port; the reference itself cannot be
imported here) on the host cores: full-depth steady-state chunks (28 layers, 32 ViT blocks, all decode forwards),
nothing extrapolated (SVLM_CPU_SAMPLE=scaled restores the layer-count-scaled sample for small hosts).
"""
import argparse
import gc
import copy
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC, UNIT = "steady_state_frames_per_s", "frames/s"
FRAMES_PER_CHUNK = 2
STEADY_ROUNDS = 200
NEW_TOKENS = 20


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:        # nvidia-smi takes a moment to print its first sample
                time.sleep(0.02)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(int(r[0]) for r in rows), "sm_max_mhz": int(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_chunk_sample(n_threads, steps=1, warmup=0, budget_s=240.0, scaled=None):
    """The reference path (oracle port) on the host cores: steady-state chunks of the same workload.  Default: FULL depth (all 28
    LLM layers, all 32 ViT blocks, every decode forward of the turn), nothing extrapolated.  `scaled` (SVLM_CPU_SAMPLE=scaled):
    4/28 layers, 8/32 blocks, 2 decode forwards, scaled by layer count.  Returns (frames/s, description, per-step chunk seconds)."""
    from oracle.cache import ListCache
    from oracle.model import Session
    from streaming_vlm_b200 import synthetic
    from streaming_vlm_b200.config import QWEN25VL_7B as G7
    if scaled is None:
        scaled = os.environ.get("SVLM_CPU_SAMPLE", "full") == "scaled"
    llm_layers, vit_blocks, decode_tokens = (4, 8, 2) if scaled else (G7.num_layers, G7.vit_depth, NEW_TOKENS - 1)
    torch.set_num_threads(n_threads)
    geo = G7.as_dict()
    geo["num_layers"], geo["vit_depth"] = llm_layers, vit_blocks
    if scaled:
        geo["fullatt_block_indexes"] = [vit_blocks - 1]            # 1 full-attention block in 8, as 4 in 32
    w = synthetic.block_weights_cpu(G7, seed=0, layers=llm_layers, vit_blocks=vit_blocks)
    state = synthetic.fast_forward(STEADY_ROUNDS, new_tokens=NEW_TOKENS)
    sess = Session(w, geo, max_new_tokens=1 + decode_tokens, suppress=synthetic.suppress_mask(G7.vocab))
    sess.prev_ids = torch.tensor([state["prev_ids"]], dtype=torch.long)
    sess.history, sess.clips, sess.pixels = state["history"], state["clips"], state["pixels"]
    sess.grids, sess.i = list(state["grids"]), state["i"]
    n = state["cache_len"]
    g = torch.Generator().manual_seed(1)
    sess.cache = ListCache()
    for _ in range(llm_layers):
        sess.cache.key_cache.append((torch.randn(1, G7.num_kv_heads, n, 128, generator=g) * 0.5).to(torch.bfloat16))
        sess.cache.value_cache.append((torch.randn(1, G7.num_kv_heads, n, 128, generator=g) * 0.5).to(torch.bfloat16))
    times = []
    t_start = time.perf_counter()
    for s in range(warmup + steps):
        i = sess.i
        px, grid = synthetic.patchify(synthetic.frames(i), G7, dtype=torch.bfloat16)
        r = sess.step(synthetic.round_prompt(i), px, grid)
        tm = sess.timing[-1]
        if scaled:
            lscale = G7.num_layers / llm_layers
            vit = (tm["vit_total"] - tm["vit_blocks"]) + tm["vit_blocks"] * (G7.vit_depth / vit_blocks)
            prefill = tm["prefill_layers"] * lscale
            dec_layers = tm["decode_layers"] / max(tm["n_decode_layers"], 1) * lscale
            # non-layer parts of a forward (embedding, positions, final norm, lm_head): GEN minus the timed layer parts
            other = max(tm["GEN"] - tm["vit_total"] - tm["prefill_layers"] - tm["decode_layers"], 0.0) / tm["n_forward"]
            chunk = tm["PKV"] * lscale + vit + prefill + other + (NEW_TOKENS - 1) * (dec_layers + other)
        else:
            chunk = tm["PKV"] + tm["GEN"]           # the reference's own sections (inference.py:312-329, 422-452), measured in full
        if s >= warmup:
            times.append(chunk)
        # pad this round's generated text to NEW_TOKENS so that the next round sees a steady-state sequence
        seq = sess.prev_ids.flatten().tolist()
        pad = NEW_TOKENS - (len(r["new_tokens"]) - 1)
        if pad > 0:
            seq = seq[:-1] + synthetic._text_tokens(pad, 991 + i) + [seq[-1]]
            sess.prev_ids = torch.tensor([seq], dtype=torch.long)
            for li in range(llm_layers):
                sess.cache.key_cache[li] = torch.cat([sess.cache.key_cache[li], sess.cache.key_cache[li][:, :, -pad:]], dim=2)
                sess.cache.value_cache[li] = torch.cat([sess.cache.value_cache[li], sess.cache.value_cache[li][:, :, -pad:]], dim=2)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    chunk = statistics.median(times)
    if scaled:
        desc = (f"1 steady-state chunk (S~{n}): eviction + ViT + prefill + {decode_tokens} of {NEW_TOKENS - 1} decode forwards, timed on "
                f"{llm_layers}/{G7.num_layers} LLM layers and {vit_blocks}/{G7.vit_depth} ViT blocks (1 full-attention) and scaled by layer "
                f"count; embedding/merger/lm_head full size; {len(times)} timed step(s), median")
    else:
        desc = (f"{len(times)} full steady-state chunk(s) (S~{n}), median: eviction + ViT (32 blocks) + prefill (28 layers) + {decode_tokens} decode "
                f"forwards, every layer executed, nothing extrapolated; PKV + GEN sections of the reference's loop")
    return FRAMES_PER_CHUNK / chunk, desc, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    # ~7 s per full chunk on 16 cores: the whole --steps K --warmup W run must end within a few minutes
    value, desc, times = cpu_chunk_sample(cores, steps=steps, warmup=min(args.warmup, 1), budget_s=200.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_streams": 1,
            "note": "ONE stream on all host cores, on rank 0, whatever --gpus says: N CPU streams would share the same cores, so "
                    "this is also the host's aggregate rate; compare it with the N = 1 line only"}
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "configs[1]: Qwen2.5-VL-7B-architecture random-init, single stream per GPU, 2 fps 448x448 synthetic video, "
                        "steady state of the sink512/text-window512/vision-window16 policy (round 200, S~5.8k), per-second "
                        "commentary decode of 20 tokens",
            "api": "reference loop body: process_past_kv + convert_qwen2_5_to_streaming(model).generate (value and e2e); "
                   "engine_api = StreamingEngine.step",
            "generation": "do_sample=True, temperature=0.9, repetition_penalty=1.05 (the reference's defaults, inference.py:41-42) for "
                          "value / e2e; greedy (the parity mode) for engine_api and the CPU arm",
            "chunk": {"frames": FRAMES_PER_CHUNK, "vit_tokens": 1024, "llm_vision_tokens": 256, "new_tokens": NEW_TOKENS,
                      "input": "uint8 RGB frames [2,448,448,3]; normalise + patchify (svlm_frames_to_patches) inside the step"},
            "policy": {"window_size": 16, "text_round": 16, "text_sink": 512, "text_sliding_window": 512},
            "cache_flush": "inputs larger than L2 (15 GB of weights and a 0.33 GB KV working set are swept every step)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def kernel_rooflines(engine, pk):
    """Per-kernel achieved throughput, CUDA events around a CUDA-graph replay of repeated launches (so that host
    launch gaps are not measured).  Algorithmic work per launch as defined in SURVEY.md section 8(d) / DESIGN.md."""
    from streaming_vlm_b200 import _abi
    g, c = engine.geo, engine.cache
    dev = engine.device
    out = {}

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        gr.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return statistics.median(ts)

    n = c.seq_len
    hq, d, L = g.num_heads, g.head_dim, g.num_layers
    # decode attention: one launch pair (split kernel + merge) per layer, N*2048 B per layer for 7B
    q1 = torch.randn(1, hq, d, device=dev).to(torch.bfloat16)
    o1 = torch.empty_like(q1)
    engine.kv_after_dev.fill_(n)
    dsplits = getattr(engine, "_splits_now", engine.decode_splits)

    def dec():
        for li in range(L):
            _abi.attn_causal(c.view, c.cur, li, q1, o1, 1, hq, c.capacity, engine.scale, dsplits, engine.ws,
                             kv_len_dev=engine.kv_after_dev)
    t = timed(dec) / L
    byt = n * 2 * g.num_kv_heads * d * 2
    out["attn_decode(split-KV+merge)"] = {"bound": "hbm", "achieved": byt / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                          "us_per_launch": t * 1e6, "launches_per_step": L * (NEW_TOKENS - 1), "kv_len": n}
    # prefill attention
    tq = 300
    qp = torch.randn(tq, hq, d, device=dev).to(torch.bfloat16)
    op = torch.empty_like(qp)
    sp = engine.prefill_splits or 2

    def pre():
        for li in range(L):
            _abi.attn_causal(c.view, c.cur, li, qp, op, tq, hq, n, engine.scale, sp, engine.ws)
    t = timed(pre) / L
    fl = 4.0 * tq * (n - tq / 2) * d * hq
    out["attn_prefill"] = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                           "us_per_launch": t * 1e6, "launches_per_step": L, "tq": tq, "kv_len": n, "splits": sp}
    # compaction + RoPE (identity gather of the n live rows; ping-pong twice per replay)
    def comp():
        _abi.kv_compact_rope(c.view, c.cur, [(0, 0, n)], c.pos, c.cos_table, c.sin_table, c.section)
        _abi.kv_compact_rope(c.view, c.cur ^ 1, [(0, 0, n)], c.pos, c.cos_table, c.sin_table, c.section)
    t = timed(comp) / 2
    byt = 2.5 * n * g.kv_bytes_per_token     # read K_raw+V, write K_raw+K_rot+V
    out["kv_compact_rope"] = {"bound": "hbm", "achieved": byt / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "us_per_launch": t * 1e6, "launches_per_step": 1, "tokens": n}
    # ViT attention (windowed, 16 windows of 64) and patch merger GEMMs
    heads, hd = g.vit_heads, g.vit_hidden // g.vit_heads
    qkv = torch.randn(1024, 3, heads, hd, device=dev).to(torch.bfloat16)
    ov = torch.empty(1024, heads, hd, dtype=torch.bfloat16, device=dev)
    cu_w = torch.arange(0, 1025, 64, dtype=torch.int32, device=dev)
    cu_f = torch.tensor([0, 1024], dtype=torch.int32, device=dev)
    t = timed(lambda: [_abi.vit_attn(qkv, cu_w, ov, 1024, heads, hd, hd ** -0.5) for _ in range(8)]) / 8
    out["vit_attn_window"] = {"bound": "tensor", "achieved": 4.0 * 16 * 64 * 64 * hd * heads / t / 1e12, "peak": pk["bf16_burst"],
                              "unit": "TFLOP/s", "us_per_launch": t * 1e6, "launches_per_step": g.vit_depth - len(g.fullatt_block_indexes)}
    t = timed(lambda: [_abi.vit_attn(qkv, cu_f, ov, 1024, heads, hd, hd ** -0.5) for _ in range(8)]) / 8
    out["vit_attn_full"] = {"bound": "tensor", "achieved": 4.0 * 1024 * 1024 * hd * heads / t / 1e12, "peak": pk["bf16_burst"],
                            "unit": "TFLOP/s", "us_per_launch": t * 1e6, "launches_per_step": len(g.fullatt_block_indexes)}
    a = torch.randn(256, 5120, device=dev).to(torch.bfloat16)
    y = torch.empty(256, 5120, dtype=torch.bfloat16, device=dev)
    t = timed(lambda: [_abi.gemm_bf16(a, engine.merger["w0"], engine.merger["b0"], y, 1) for _ in range(8)]) / 8
    out["patch_merger_gemm0"] = {"bound": "tensor", "achieved": 2.0 * 256 * 5120 * 5120 / t / 1e12, "peak": pk["bf16_burst"],
                                 "unit": "TFLOP/s", "us_per_launch": t * 1e6, "launches_per_step": 1}
    # decode GEMVs (weight streaming, one launch per projection per layer; 28 layers = 13 GB >> L2)
    x = torch.randn(g.hidden, device=dev).to(torch.bfloat16)
    xr = torch.randn(g.hidden, device=dev).to(torch.bfloat16)
    qkv_o = torch.empty((hq + 2 * g.num_kv_heads) * d, dtype=torch.bfloat16, device=dev)
    att = torch.randn(hq * d, device=dev).to(torch.bfloat16)
    act = torch.empty(g.mlp, dtype=torch.bfloat16, device=dev)
    for name, fn, nbytes in [
        ("gemv_norm_qkv_rope_append", lambda Lw: _abi.gemv_qkv_rope_append(c.view, c.cur, 0, Lw["wqkv"], x, Lw["ln1"], g.rms_eps, Lw["bqkv"], hq, q1,
                                                                             engine.kv_len_dev, c.next_pos, c.cos_table, c.sin_table),
         engine.layers[0]["wqkv"].numel() * 2),
        ("gemv_o_residual", lambda Lw: _abi.gemv_bf16(Lw["wo"], att, xr, 1, residual=xr), engine.layers[0]["wo"].numel() * 2),
        ("gemv_norm_gate_up_swiglu", lambda Lw: _abi.gemv_bf16(Lw["wgu"], x, act, 2, Lw["ln2"], g.rms_eps), engine.layers[0]["wgu"].numel() * 2),
        ("gemv_down_residual", lambda Lw: _abi.gemv_bf16(Lw["wd"], act, xr, 1, residual=xr), engine.layers[0]["wd"].numel() * 2),
    ]:
        t = timed(lambda: [fn(Lw) for Lw in engine.layers]) / L
        out[name] = {"bound": "hbm", "achieved": nbytes / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "us_per_launch": t * 1e6,
                     "launches_per_step": L * (NEW_TOKENS - 1), "weight_bytes": nbytes}
    key = torch.zeros(1, dtype=torch.int64, device=dev)
    tok = torch.zeros(1, dtype=torch.int64, device=dev)
    cnt = torch.zeros(3, dtype=torch.int32, device=dev)
    t = timed(lambda: [_abi.lm_head_argmax(engine.lm_head, x, engine.norm, g.rms_eps, engine.suppress_u8, key, tok, cnt[0:1], cnt[1:2], cnt[2:3])
                       for _ in range(4)]) / 4
    out["gemv_norm_lm_head_argmax"] = {"bound": "hbm", "achieved": engine.lm_head.numel() * 2 / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                       "us_per_launch": t * 1e6, "launches_per_step": NEW_TOKENS - 1, "weight_bytes": engine.lm_head.numel() * 2}
    traffic = {}
    for name in ("r2_traffic.json", "r1_traffic.json"):          # DRAM bytes per launch from the committed ncu --set full captures
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["bytes_per_launch"]
            break
        except (OSError, KeyError, ValueError):
            continue
    for k, v in out.items():
        v["frac"] = v["achieved"] / v["peak"]
        v["traffic"] = traffic.get(k)          # DRAM bytes per launch from the committed ncu --set full capture (or null)
        v["share_ms_per_step"] = v["us_per_launch"] * v["launches_per_step"] * 1e-3
    return out


def multi_stream_leg(engine, n_streams, rounds, warmup, rank, dev, pk):
    """BASELINE configs[3] at this GPU: `n_streams` independent 7B streams sharing the GPU (and one copy of the weights), each
    in the same steady state as the headline stream but with its own frames, prompts and cache contents.  One ROUND = one
    chunk of every stream (per-stream eviction, batched ViT + chunk prefill, then batched decode steps).  Pixel values resident in HBM."""
    from streaming_vlm_b200 import _abi, synthetic
    from streaming_vlm_b200.config import QWEN25VL_7B as G7
    from streaming_vlm_b200.multi_stream import MultiStreamEngine
    multi = MultiStreamEngine(None, G7, n_streams, capacity=8192, device=dev, suppress=synthetic.suppress_mask(G7.vocab),
                              share_weights_from=engine)
    total = warmup + rounds
    prompts, pixels = [], []
    grid = None
    for b, e in enumerate(multi.engines):
        seed = 1000 * rank + 100 + b
        e.load_state(synthetic.fast_forward(STEADY_ROUNDS, new_tokens=NEW_TOKENS, seed=seed), seed=seed)
    base_i = multi.engines[0].i
    for k in range(total):
        pr, px = [], []
        for b in range(n_streams):
            p_, grid = synthetic.patchify(synthetic.frames(7919 * (b + 1) + base_i + k), G7, dtype=torch.float32)
            px.append(p_.to(dev))
            pr.append(synthetic.round_prompt(base_i + k, seed=1000 * rank + 100 + b))
        prompts.append(pr)
        pixels.append(px)
    toks = 0
    for k in range(warmup):
        multi.step(prompts[k], pixels[k], grid, sync_timing=False)
    torch.cuda.synchronize()
    l0 = multi.total_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_fwd = 0
    for k in range(warmup, total):
        res = multi.step(prompts[k], pixels[k], grid, sync_timing=False)
        toks += sum(min(len(r["new_tokens"]), NEW_TOKENS) for r in res)
        n_fwd += multi.timing[-1]["n_forward"] - 1
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    out = {"streams": n_streams, "value": n_streams * rounds * FRAMES_PER_CHUNK / t, "unit": UNIT, "rounds": rounds, "warmup": warmup,
           "ms_per_round": 1e3 * t / rounds, "decode_tok_s": toks / t, "kv_len": [e.cache.seq_len for e in multi.engines],
           "gpu_launches": multi.total_launches() - l0, "batched_decode_steps": n_fwd,
           "note": "per-stream eviction; ViT and chunk prefill of all streams in one pass; ONE batched decode step per token for all streams (one weight sweep each)"}
    # the batched split-KV decode attention on its own: one launch pair per layer over all streams' caches
    g, L = G7, G7.num_layers
    caches = [e.cache for e in multi.engines]
    slots = _abi.make_slots(caches, list(multi.kv_len.split(1)), list(multi.next_pos.split(1)), list(multi.kv_after.split(1)))
    for b, c in enumerate(caches):
        multi.kv_after[b] = c.seq_len
    q = torch.randn(n_streams, g.num_heads, g.head_dim, device=dev).to(torch.bfloat16)
    o = torch.empty_like(q)
    n_tiles = max(c.seq_len for c in caches) // 128 + 1
    splits = max(1, min(multi.max_splits, -(-n_tiles // 3)))

    def att():
        for li in range(L):
            _abi.attn_decode_streams(caches[0].view, slots, li, q, o, g.num_heads, caches[0].capacity, multi.engines[0].scale, splits, multi.ws)
    att()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        att()
    ts = []
    for _ in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); gr.replay(); a1.record(); torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1) * 1e-3 / L)
    ta = statistics.median(ts)
    byt = sum(c.seq_len for c in caches) * 2 * g.num_kv_heads * g.head_dim * 2
    out["attn_decode_streams"] = {"bound": "hbm", "achieved": byt / ta / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": byt / ta / 1e9 / pk["hbm_gbs"],
                                  "us_per_launch": ta * 1e6, "splits": splits, "alg_bytes_per_launch": byt}
    # the projections of the batched decode step: svlm_gemm_tokens, one launch each, 28 distinct weight sets per replay (>> L2)
    e0 = multi.engines[0]
    B = n_streams
    xs = torch.randn(B, g.hidden, device=dev).to(torch.bfloat16)
    xr = torch.randn(B, g.hidden, device=dev).to(torch.bfloat16)
    att = torch.randn(B, g.num_heads * g.head_dim, device=dev).to(torch.bfloat16)
    actb = torch.randn(B, g.mlp, device=dev).to(torch.bfloat16)
    stat = torch.rand(-(-g.hidden // 128), B, device=dev) * 128
    qrot = torch.empty(B, g.num_heads, g.head_dim, dtype=torch.bfloat16, device=dev)
    c0 = caches[0]
    gt = {}
    for name, fn, nbytes in [
        ("gemm_tokens_norm_qkv_rope_append", lambda li, Lw: _abi.gemm_tokens_qkv_append(c0.view, slots, li, xs, Lw["wqkv"], Lw["bqkv"], Lw["ln1"], g.rms_eps,
                                                                                      g.num_heads, qrot, c0.cos_table, c0.sin_table, stats_in=stat),
         e0.layers[0]["wqkv"].numel() * 2),
        ("gemm_tokens_o_residual", lambda li, Lw: _abi.gemm_tokens(att, Lw["wo"], xr, residual=xr, epilogue=_abi.EPI_RESIDUAL, stats_out=stat),
         e0.layers[0]["wo"].numel() * 2),
        ("gemm_tokens_norm_gate_up_swiglu", lambda li, Lw: _abi.gemm_tokens(xs, Lw["wgu"], actb, epilogue=_abi.EPI_SWIGLU, norm_weight=Lw["ln2"],
                                                                            eps=g.rms_eps, stats_in=stat), e0.layers[0]["wgu"].numel() * 2),
        ("gemm_tokens_down_residual", lambda li, Lw: _abi.gemm_tokens(actb, Lw["wd"], xr, residual=xr, epilogue=_abi.EPI_RESIDUAL, stats_out=stat),
         e0.layers[0]["wd"].numel() * 2),
    ]:
        def run(fn=fn):
            for li, Lw in enumerate(e0.layers):
                fn(li, Lw)
        run()
        torch.cuda.synchronize()
        grk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(grk):
            run()
        grk.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); grk.replay(); a1.record(); torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1) * 1e-3 / L)
        tk = statistics.median(ts)
        gt[name] = {"bound": "hbm", "achieved": nbytes / tk / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": nbytes / tk / 1e9 / pk["hbm_gbs"],
                    "us_per_launch": tk * 1e6, "tokens": B, "weight_bytes": nbytes}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        traffic = {}
    for k_, v_ in gt.items():
        v_["traffic"] = traffic.get(k_)
    out["decode_projections"] = gt
    out["attn_decode_streams"]["traffic"] = traffic.get("attn_decode_streams") if n_streams == 8 else None
    # the chunk-prefill attention of all streams the way the batched prefill runs it: one un-split launch per stream, the streams'
    # launches side by side on their own CUDA streams (fork / join inside the graph), Tq = 300 new tokens each
    tq = 300
    qp = torch.randn(n_streams * tq, g.num_heads, g.head_dim, device=dev).to(torch.bfloat16)
    op = torch.empty_like(qp)

    def pre():
        main = torch.cuda.current_stream()
        for li in range(L):
            for b, c in enumerate(caches):
                side = multi._pf_streams[b]
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    _abi.attn_causal(c.view, c.cur, li, qp[b * tq:(b + 1) * tq], op[b * tq:(b + 1) * tq], tq, g.num_heads, c.seq_len, multi.engines[0].scale, 1, None)
            for b in range(n_streams):
                main.wait_stream(multi._pf_streams[b])
    pre()
    torch.cuda.synchronize()
    gr2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr2):
        pre()
    gr2.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); gr2.replay(); a1.record(); torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1) * 1e-3 / L)
    tp_ = statistics.median(ts)
    fl = sum(4.0 * tq * (c.seq_len - tq / 2) * g.head_dim * g.num_heads for c in caches)
    out["attn_prefill_streams"] = {"bound": "tensor", "achieved": fl / tp_ / 1e12, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": fl / tp_ / 1e12 / pk["bf16_burst"],
                                   "us_per_layer_all_streams": tp_ * 1e6, "tq": tq, "alg_flop_per_layer": fl,
                                   "note": "8 un-split launches per layer, one per stream, concurrent on 8 CUDA streams"}
    return out


def build_hf_model(geo, weights, dev):
    """A stock transformers Qwen2_5_VLForConditionalGeneration of geometry `geo` whose parameters ARE the tensors of `weights`
    (constructed on the meta device, then `load_state_dict(assign=True)`: no second copy, no init pass over 8 B parameters)."""
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    text = dict(hidden_size=geo.hidden, num_hidden_layers=geo.num_layers, num_attention_heads=geo.num_heads, num_key_value_heads=geo.num_kv_heads,
                intermediate_size=geo.mlp, vocab_size=geo.vocab, rms_norm_eps=geo.rms_eps, tie_word_embeddings=False,
                rope_parameters={"rope_type": "default", "rope_theta": geo.rope_theta, "mrope_section": list(geo.mrope_section)})
    vis = dict(depth=geo.vit_depth, hidden_size=geo.vit_hidden, num_heads=geo.vit_heads, intermediate_size=geo.vit_mlp, out_hidden_size=geo.hidden,
               patch_size=geo.patch_size, temporal_patch_size=geo.temporal_patch_size, spatial_merge_size=geo.spatial_merge_size,
               window_size=geo.window_size, fullatt_block_indexes=list(geo.fullatt_block_indexes), tokens_per_second=geo.tokens_per_second)
    cfg = Qwen2_5_VLConfig(text_config=text, vision_config=vis, tie_word_embeddings=False)
    with torch.device("meta"):
        model = Qwen2_5_VLForConditionalGeneration(cfg).to(torch.bfloat16)
    missing, unexpected = model.load_state_dict(weights, assign=True, strict=False)
    if missing or unexpected:
        raise RuntimeError(f"synthetic weights do not cover the HF model: missing {missing[:4]}, unexpected {unexpected[:4]}")
    return model.eval()


class DropInSession:
    """The reference's per-stream loop state and loop body (`inference/inference.py:297-303, 309-517`), product side: everything a
    maintainer's unchanged `streaming_inference` does between reading two frames and appending the assistant turn."""

    def __init__(self, model, geo, state, dev, sup, gen_kwargs):
        from streaming_vlm_b200.inference.streaming_args import StreamingArgs
        self.model, self.geo, self.dev, self.sup, self.gen_kwargs = model, geo, dev, sup, gen_kwargs
        self.args = StreamingArgs(pos_mode="shrink")
        self.past = model._svlm_engine.cache
        self.prev = torch.tensor([state["prev_ids"]], dtype=torch.long, device=dev)
        self.history, self.clips, self.pixels = state["history"], state["clips"], state["pixels"]
        self.i = int(state["i"])
        self.args.video_grid_thw = torch.tensor([list(g) for g in state["grids"]], dtype=torch.long, device=dev)
        self.timing = []

    def step(self, prompt_ids, frames_thwc, max_new_tokens=None):
        from streaming_vlm_b200.inference.inference import process_past_kv
        max_new_tokens = NEW_TOKENS if max_new_tokens is None else int(max_new_tokens)
        torch.ops.svlm  # noqa: B018  (custom ops registered by convert_qwen2_5_to_streaming)
        i = self.i
        t0 = time.perf_counter()
        self.past, self.prev, self.clips, self.pixels = process_past_kv(
            self.past, i, DEFAULT_POLICY["text_round"], DEFAULT_POLICY["window_size"], self.history, self.prev, 3, 2, self.clips, self.pixels,
            DEFAULT_POLICY["text_sink"], DEFAULT_POLICY["text_sliding_window"])
        t_pkv = time.perf_counter() - t0
        self.clips.append(None)
        self.history.append({"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}]})
        # the pixel half of `processor(text=..., videos=clip)` (inference.py:390-395): normalise + patchify on the GPU
        px = torch.ops.svlm.frames_to_patches(frames_thwc, self.geo.patch_size, self.geo.temporal_patch_size, self.geo.spatial_merge_size)
        grid = torch.tensor([[frames_thwc.shape[0] // self.geo.temporal_patch_size, frames_thwc.shape[1] // self.geo.patch_size,
                              frames_thwc.shape[2] // self.geo.patch_size]], dtype=torch.long)
        ids = torch.tensor([prompt_ids], dtype=torch.long, device=self.dev)
        ids = torch.cat([self.prev, ids if self.prev[0, -1].item() != 198 else ids[:, 1:]], dim=1)          # inference.py:397-406
        self.pixels.append(None)
        self.args.input_ids = ids
        self.args.video_grid_thw = torch.cat([self.args.video_grid_thw, grid.to(self.dev)], dim=0)         # grows, as in the reference (:415)
        n_in = ids.shape[1]
        t0 = time.perf_counter()
        out = self.model.generate(input_ids=ids, pixel_values_videos=px, video_grid_thw=grid, past_key_values=self.past,
                                  max_new_tokens=max_new_tokens, use_cache=True, return_dict_in_generate=True, streaming_args=self.args,
                                  pad_token_id=151645, suppress_tokens=self.sup, **self.gen_kwargs)
        t_gen = time.perf_counter() - t0
        seq = out.sequences
        if seq[0, -1].item() != 151645:                                                                    # inference.py:457-459
            seq = torch.cat([seq, torch.tensor([[151645]], device=self.dev)], dim=1)
        self.past = out.past_key_values
        self.prev = seq.clone()
        new = seq.shape[1] - n_in
        self.history.append({"role": "assistant", "content": "x " * new + "..."})
        self.timing.append({"PKV": t_pkv, "GEN": t_gen, "new_tokens": new})
        self.i += 1
        return new


DEFAULT_POLICY = {"window_size": 16, "text_round": 16, "text_sink": 512, "text_sliding_window": 512}


def config3_sweep(pk, dev, sizes=(4096, 8192, 16384, 32768, 65536)):
    """BASELINE.json configs[2]: KV compaction + contiguous-RoPE, split-KV decode attention and chunk-prefill attention at retained
    cache lengths 4k..64k (7B geometry, bf16), each as a fraction of its roofline, under the same clock sampler as the headline.
    Split counts are the engine's own heuristics.  (benchmarks/microbench.py adds the reference's library calls on the same GPU.)"""
    from streaming_vlm_b200 import _abi
    from streaming_vlm_b200.config import QWEN25VL_7B as G
    from streaming_vlm_b200.engine import decode_split_count
    from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache
    L, H, HQ, D = G.num_layers, G.num_kv_heads, G.num_heads, G.head_dim
    scale = 1.0 / math.sqrt(D)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        gr.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return statistics.median(ts)

    points = []
    with ClockSampler(dev.index or 0) as clk:
        for n in sizes:
            cap = n + 512
            c = StreamingCache(G, cap, dev)
            c.arena[0].normal_(0, 0.5)
            c.pos.copy_(torch.arange(cap, dtype=torch.int32, device=dev).expand(3, cap) // 4)
            # the steady-state 5-run gather scaled to n: sink 527 | drop 10 | text window | moved 10 | 2 | drop ~287 | rest
            n_src = n + 297
            runs, d = [], 0
            for s_, l in [(0, 527), (537, 500), (1314, 10), (1037, 2), (1327, n_src - 1327)]:
                runs.append((s_, d, l))
                d += l
            keep = d
            rec = {"retained_tokens": keep}
            t = timed(lambda: _abi.kv_compact_rope(c.view, 0, runs, c.pos, c.cos_table, c.sin_table, c.section))
            byt = 2.5 * keep * G.kv_bytes_per_token
            rec["kv_compact_rope"] = {"us": t * 1e6, "GBs": byt / t / 1e9, "frac": byt / t / 1e9 / pk["hbm_gbs"], "bound": "hbm"}
            q1 = torch.randn(1, HQ, D, device=dev).to(torch.bfloat16)
            o1 = torch.empty_like(q1)
            kv_dev = torch.tensor([keep], dtype=torch.int32, device=dev)
            splits = decode_split_count(keep, H)                                  # StreamingEngine.decode_step
            ws = torch.zeros(max(_abi.attn_workspace_bytes(1, HQ, D, splits), _abi.attn_workspace_bytes(512, HQ, D, 4)) // 4 + 16,
                             dtype=torch.float32, device=dev)
            t = timed(lambda: [_abi.attn_causal(c.view, 1, li, q1, o1, 1, HQ, cap, scale, splits, ws, kv_len_dev=kv_dev) for li in range(L)]) / L
            byt = keep * 2 * H * D * 2
            rec["attn_decode"] = {"us_per_layer": t * 1e6, "splits": splits, "GBs": byt / t / 1e9, "frac": byt / t / 1e9 / pk["hbm_gbs"], "bound": "hbm"}
            tq = 276
            qp = torch.randn(tq, HQ, D, device=dev).to(torch.bfloat16)
            op = torch.empty_like(qp)
            sp = 1 if (-(-tq // (128 // G.group))) * H >= 74 or keep < 2048 else 2     # StreamingEngine.prefill
            t = timed(lambda: [_abi.attn_causal(c.view, 1, li, qp, op, tq, HQ, keep, scale, sp, ws) for li in range(L)]) / L
            fl = 4.0 * tq * (keep - tq / 2) * D * HQ
            rec["attn_prefill"] = {"us_per_layer": t * 1e6, "splits": sp, "tq": tq, "TFLOPs": fl / t / 1e12, "frac": fl / t / 1e12 / pk["bf16_burst"],
                                   "bound": "tensor"}
            points.append(rec)
            del c, ws
            torch.cuda.empty_cache()
    return {"workload": "configs[2]: retained cache 4k-64k tokens, 7B geometry (28 layers, 4 kv heads, 28 q heads, d 128), bf16; CUDA events "
                        "around CUDA-graph replays of 28 layer launches (K/V of 28 layers >> L2 from 8k tokens up)",
            "peaks": {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops_burst": pk["bf16_burst"]}, "clocks": clk.summary(), "points": points}


def run_gpu(args):
    from streaming_vlm_b200 import _abi, synthetic
    from streaming_vlm_b200.config import QWEN25VL_7B as G7
    from streaming_vlm_b200.engine import GREEDY
    from streaming_vlm_b200.inference.generate.streaming_generate_qwen import _fast_path_engine
    from streaming_vlm_b200.inference.qwen2_5.patch_model import convert_qwen2_5_to_streaming
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _abi.build()
    dev = torch.device("cuda", local)
    pk = peaks()

    # ---- the reference-facing object: a stock HF model, patched; its `generate` runs on the session runner over the SAME tensors
    w = synthetic.random_weights(G7, seed=rank, device=dev)
    model = convert_qwen2_5_to_streaming(build_hf_model(G7, w, dev))
    del w
    engine = _fast_path_engine(model, None, 8192)
    sup = synthetic.suppress_mask(G7.vocab).to(dev)
    engine.set_suppress(sup)
    state = synthetic.fast_forward(STEADY_ROUNDS, new_tokens=NEW_TOKENS, seed=rank)
    engine.load_state(state, seed=rank)
    gen_kwargs = dict(do_sample=True, temperature=0.9, repetition_penalty=1.05) if args.generation == "reference" else dict(do_sample=False)
    sess = DropInSession(model, G7, state, dev, sup, gen_kwargs)
    n_legs = 3
    total = args.warmup + n_legs * args.steps + 2
    base_i = engine.i
    # on-wire format of a chunk: 2 RGB frames, uint8 [T, H, W, 3] (what a video decoder hands over); the ingest kernel
    # (normalise + patchify, svlm_frames_to_patches) runs inside every timed step
    fr_host = [synthetic.frames(base_i + k).permute(0, 2, 3, 1).contiguous().pin_memory() for k in range(total)]
    fr_dev = [f.to(dev) for f in fr_host[:args.warmup + args.steps]]
    prompts = [synthetic.round_prompt(base_i + k, seed=rank) for k in range(total)]
    grid = (1, fr_host[0].shape[1] // G7.patch_size, fr_host[0].shape[2] // G7.patch_size)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_dropin(k0, n, host_pixels):
        lat, toks = [], 0
        for k in range(k0, k0 + n):
            t0 = time.perf_counter()
            fr = fr_host[k].to(dev, non_blocking=True) if host_pixels else fr_dev[k]
            toks += min(sess.step(prompts[k], fr), NEW_TOKENS)
            torch.cuda.current_stream().synchronize()
            lat.append(time.perf_counter() - t0)
        return lat, toks

    run_dropin(0, args.warmup, False)
    # ---- timed region 1 (value): drop-in API, frames resident in HBM
    barrier()
    launches0 = engine.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        if args.profile_range:
            torch.cuda.profiler.start()
        e0.record()
        lat, toks = run_dropin(args.warmup, args.steps, False)
        e1.record()
        barrier()
        if args.profile_range:
            torch.cuda.profiler.stop()
        t_dev = e0.elapsed_time(e1) * 1e-3
        launches = engine.launches - launches0
        gen_s = sum(t["GEN"] for t in sess.timing[-args.steps:])
        gen_tok = sum(t["new_tokens"] for t in sess.timing[-args.steps:])
        # ---- timed region 2 (e2e): drop-in API, uint8 frames from pinned host memory, ids up / tokens down through the host
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        lat2, toks2 = run_dropin(args.warmup + args.steps, args.steps, True)
        f1.record()
        barrier()
        t_e2e = f0.elapsed_time(f1) * 1e-3
    n_ids = int(sess.prev.shape[1])
    # ---- the same stream continued through StreamingEngine.step directly (greedy): what the drop-in costs on top of the engine
    engine.prev_ids = sess.prev[0].cpu().numpy()
    engine.history, engine.clips, engine.pixels, engine.i = sess.history, sess.clips, sess.pixels, sess.i
    engine.grids = [tuple(grid)] * (len(state["grids"]) + sess.i - base_i)
    engine._grids_uploaded = None
    engine.gen = GREEDY
    k_eng = args.warmup + 2 * args.steps
    engine.step(prompts[k_eng], None, grid, sync_timing=False, frames=fr_host[k_eng].to(dev))          # captures the greedy graphs
    engine.step(prompts[k_eng + 1], None, grid, sync_timing=False, frames=fr_host[k_eng + 1].to(dev))
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    toks3 = 0
    for k in range(k_eng + 2, k_eng + 2 + args.steps):
        r = engine.step(prompts[k], None, grid, sync_timing=False, frames=fr_host[k].to(dev, non_blocking=True))
        toks3 += min(len(r["new_tokens"]), NEW_TOKENS)
    g1.record()
    barrier()
    t_eng = g0.elapsed_time(g1) * 1e-3
    if dist is not None:
        tt = torch.tensor([t_dev, t_e2e, t_eng], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_eng = float(tt[0]), float(tt[1]), float(tt[2])

    h2d = fr_host[0].numel() + n_ids * 8            # uint8 frames + the int64 ids the loop uploads (torch.tensor(prompt), pruned ids)
    d2h = n_ids * 8 + NEW_TOKENS * 8                # the loop reads the ids back for eviction planning, plus one id per token
    line = None
    if rank == 0:
        roofs = kernel_rooflines(engine, pk)
        dom = max(roofs, key=lambda k: roofs[k]["share_ms_per_step"])
        roof = {k: roofs[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
        roof["kernel"] = dom
        roof["peak_source"] = pk["source"] + "; HBM peak = measured COPY bandwidth (read + write), read-only streams can slightly exceed it"
        line = {"metric": METRIC, "value": world * args.steps * FRAMES_PER_CHUNK / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(),
                "decode_tok_s": world * toks / t_dev, "p50_chunk_ms": 1e3 * statistics.median(lat), "p99_chunk_ms": 1e3 * max(lat),
                # the reference's own protocol number (eval/efficiency/efficiency_test.py:88-98): GEN section / decoded tokens
                "gen_s_per_token": gen_s / max(gen_tok, 1),
                "e2e": {"value": world * args.steps * FRAMES_PER_CHUNK / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "p50_chunk_ms": 1e3 * statistics.median(lat2), "decode_tok_s": world * toks2 / t_e2e,
                        "api": "process_past_kv + patched HF model.generate (host frames, host ids)"},
                "engine_api": {"value": world * args.steps * FRAMES_PER_CHUNK / t_eng, "unit": UNIT, "ms_per_step": 1e3 * t_eng / args.steps,
                               "decode_tok_s": world * toks3 / t_eng, "generation": "greedy", "api": "StreamingEngine.step (host frames)"},
                "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "roofline_kernels": roofs,
                "kv_len": engine.cache.seq_len, "streams_per_gpu": 1,
                "graphs": {"prefill": len(engine._pf_graphs), "decode": len(engine._graphs), "vit": len(engine._vit_graphs)}}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, desc, _ = cpu_chunk_sample(cores, steps=1, warmup=0, budget_s=60.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        else:
            line["cpu_baseline"] = None
        if world == 1 and not args.no_sweep:
            try:
                line["config3_sweep"] = config3_sweep(pk, dev)
            except Exception as exc:   # noqa: BLE001   (an extra leg must never cost the headline line)
                line["config3_sweep"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    for key, n_streams in (("multi_stream", args.streams), ("multi_stream_16", args.streams2)):
        if n_streams <= 1 or (key == "multi_stream_16" and n_streams == args.streams):
            continue
        # extra legs: they must never cost the headline line (every rank still takes part in the all-reduce below)
        try:
            ms = multi_stream_leg(engine, n_streams, max(2, args.steps // 2), 2, rank, dev, pk)
        except Exception as exc:   # noqa: BLE001
            ms = {"streams": n_streams, "value": 0.0, "unit": UNIT, "error": f"{type(exc).__name__}: {exc}"[:300]}
        gc.collect()
        torch.cuda.empty_cache()
        if dist is not None:
            tt = torch.tensor([ms["value"]], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            ms["value_all_gpus"] = float(tt[0])
        if rank == 0:
            line[key] = ms
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_tp(args):
    """BASELINE.json configs[4]: ONE stream of the 32B (or 7B) architecture, its language model head-sharded over the `--tp` ranks
    of this torchrun job.  The row-parallel projections are reduced by our own kernels over NVLink symmetric memory (no NCCL on the
    data path), every step is a CUDA-graph replay.  `value` = frames/s of the one stream (strong scaling: the work is fixed)."""
    import torch.distributed as dist
    from streaming_vlm_b200 import _abi, parallel, synthetic
    from streaming_vlm_b200.config import PRESETS
    from streaming_vlm_b200.engine import StreamingEngine
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.tp:
        raise RuntimeError(f"--tp {args.tp} needs a torchrun job of {args.tp} ranks (got WORLD_SIZE = {world})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    geo = PRESETS[args.geometry]
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.ones(1, device=dev))
    _abi.build()
    w = synthetic.random_weights(geo, seed=0, device=dev)
    if world > 1:
        w, local_geo = parallel.shard_llm_weights(w, geo, rank, world)
    else:
        local_geo = geo
    torch.cuda.empty_cache()
    eng = StreamingEngine(w, local_geo, capacity=8192, device=dev, suppress=synthetic.suppress_mask(geo.vocab),
                          tp_group=dist.group.WORLD if world > 1 else None)
    del w
    torch.cuda.empty_cache()
    eng.load_state(synthetic.fast_forward(STEADY_ROUNDS, new_tokens=NEW_TOKENS, seed=0), seed=0)
    steps, warmup = args.steps, max(args.warmup, 3)
    base = eng.i
    frames = [synthetic.frames(base + k).permute(0, 2, 3, 1).contiguous().pin_memory() for k in range(steps + warmup)]
    prompts = [synthetic.round_prompt(base + k) for k in range(steps + warmup)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(warmup):
        eng.step(prompts[k], None, None, sync_timing=False, frames=frames[k].to(dev))
    barrier()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        toks = 0
        for k in range(warmup, warmup + steps):
            r = eng.step(prompts[k], None, None, sync_timing=False, frames=frames[k].to(dev, non_blocking=True))
            toks += min(len(r["new_tokens"]), NEW_TOKENS)
        e1.record()
        barrier()
    t = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    if rank == 0:
        ar = 2 * geo.num_layers
        line = {"metric": METRIC, "value": steps * FRAMES_PER_CHUNK / t, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"configs[4]: {geo.name} random-init, ONE stream, language model head-sharded over {world} GPU(s) "
                                       "(q/k/v/gate/up column-parallel, o_proj/down_proj row-parallel + one-shot all-reduce over NVLink symmetric "
                                       "memory inside svlm_tp_gemv_allreduce / svlm_tp_allreduce_rows), steady state S~5.8k, 20 tokens per chunk, greedy",
                           "tp": world, "geometry": args.geometry, "cuda_graphs": True,
                           "cache_flush": "inputs larger than L2 (the weight shard of every rank is swept every token)"},
                "decode_tok_s": toks / t, "gpu_launches": eng.launches - l0, "clocks": clk.summary(), "kv_len": eng.cache.seq_len,
                "e2e": {"value": steps * FRAMES_PER_CHUNK / t, "unit": UNIT, "h2d_bytes_per_step": frames[0].numel() + len(eng.prev_ids) * 8,
                        "d2h_bytes_per_step": NEW_TOKENS * 8, "note": "the timed region itself copies the uint8 frames from pinned host memory and reads the tokens back"},
                "all_reduces_per_token": ar if world > 1 else 0,
                "nvlink_bytes_pulled_per_token_per_rank": ar * eng.tp.nvlink_bytes_per_decode_allreduce() if world > 1 else 0,
                "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "cpu_baseline": None, "roofline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=8, help="streams per GPU of the extra multi-stream leg (BASELINE configs[3]); 1 = skip")
    ap.add_argument("--streams2", type=int, default=16, help="streams per GPU of the second multi-stream leg (`multi_stream_16`); 1 = skip")
    ap.add_argument("--generation", default="reference", choices=["reference", "greedy"],
                    help="token selection of the timed drop-in legs: the reference's sampling config or greedy")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[2] 4k-64k kernel sweep")
    ap.add_argument("--tp", type=int, default=0, help="tensor-parallel degree of ONE stream (configs[4]); the torchrun job must have that many ranks; "
                                                      "1 = the same stream unsharded on one GPU")
    ap.add_argument("--geometry", default="32b", choices=["7b", "32b"], help="model geometry of the --tp mode")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around timed region 1 (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)      # timing rule: at least 3 untimed warm-up steps (CUDA-graph capture happens there)
    if args.impl == "reference":
        run_reference(args)
    elif args.tp:
        run_tp(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
