#!/usr/bin/env python
"""BASELINE config 3: KV compaction + contiguous-RoPE + decode / prefill attention microbench, retained cache
4k..64k tokens, 7B geometry (28 layers, 4 kv heads, 28 q heads, d 128, bf16), achieved GB/s and TFLOP/s against
the measured B200 peaks, next to the reference's own library calls on the same GPU (torch.index_select / cat for
eviction, elementwise bf16 mRoPE over the whole K, repeat_kv + flash_attn_func for attention).

    python benchmarks/microbench.py > profiles/microbench_rNN.jsonl
"""
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from streaming_vlm_b200 import _abi  # noqa: E402
from streaming_vlm_b200.config import QWEN25VL_7B as G  # noqa: E402
from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache  # noqa: E402


def timed(fn, reps=7, graph=True):
    fn()
    torch.cuda.synchronize()
    if graph:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        run = gr.replay
    else:
        run = fn
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)


def main():
    from bench import ClockSampler
    with ClockSampler(0) as clk:
        sweep()
    print(json.dumps({"clocks": clk.summary()}), flush=True)


def sweep():
    _abi.build()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    dev = torch.device("cuda")
    L, H, HQ, D = G.num_layers, G.num_kv_heads, G.num_heads, G.head_dim
    scale = 1.0 / math.sqrt(D)
    try:
        from flash_attn import flash_attn_func
    except Exception:  # noqa: BLE001
        flash_attn_func = None
    for n in [4096, 8192, 16384, 32768, 65536]:
        cap = n + 512
        c = StreamingCache(G, cap, dev)
        c.arena[0].normal_(0, 0.5)
        c.seq_len = n
        c.pos.copy_(torch.arange(cap, dtype=torch.int32, device=dev).clamp_(max=cap - 1).expand(3, cap) // 4)
        # the steady-state 5-run gather scaled to n: sink 527 | drop 10 | text window | moved 10 | 2 | drop ~287 | rest
        n_src = n + 297
        runs_src = [(0, 527), (537, 500), (1314, 10), (1037, 2), (1327, n_src - 1327)]
        runs, d = [], 0
        for s, l in runs_src:
            runs.append((s, d, l))
            d += l
        keep = d
        rec = {"retained_tokens": keep, "layers": L, "kv_heads": H, "q_heads": HQ}

        def comp():
            _abi.kv_compact_rope(c.view, 0, runs, c.pos, c.cos_table, c.sin_table, c.section)
        t = timed(comp)
        byt = 2.5 * keep * G.kv_bytes_per_token
        rec["kv_compact_rope"] = {"us": t * 1e6, "GBs": byt / t / 1e9, "frac_hbm": byt / t / 1e9 / pk["hbm_gbs"], "alg_bytes": byt}

        # reference library path for the same eviction: cat + 3 index_select over every layer's K and V, then the
        # whole-K elementwise mRoPE of one forward (language_forward.py:58-66)
        kref = [torch.randn(1, H, n_src, D, device=dev).to(torch.bfloat16) for _ in range(L)]
        vref = [torch.randn(1, H, n_src, D, device=dev).to(torch.bfloat16) for _ in range(L)]
        idx = torch.cat([torch.arange(s, s + l, device=dev) for s, l in runs_src])

        def ref_evict():
            for li in range(L):
                k, v = kref[li], vref[li]
                for _ in range(3):
                    k2 = torch.index_select(k, 2, idx)
                    v2 = torch.index_select(v, 2, idx)
        t = timed(ref_evict, graph=False)
        rec["ref_torch_index_select_x3"] = {"us": t * 1e6}
        cos = torch.randn(1, 1, keep, D, device=dev).to(torch.bfloat16)
        sin = torch.randn(1, 1, keep, D, device=dev).to(torch.bfloat16)
        kk = [torch.randn(1, H, keep, D, device=dev).to(torch.bfloat16) for _ in range(L)]

        def ref_rope():
            for li in range(L):
                k = kk[li]
                h = D // 2
                rot = torch.cat((-k[..., h:], k[..., :h]), dim=-1)
                _ = (k * cos) + (rot * sin)
        t = timed(ref_rope, graph=False)
        rec["ref_torch_mrope_whole_K_per_forward"] = {"us": t * 1e6}
        del kref, vref, kk

        # decode attention over n keys, all layers back to back (graph), split-KV
        c.cur = 1
        c.seq_len = keep
        q1 = torch.randn(1, HQ, D, device=dev).to(torch.bfloat16)
        o1 = torch.empty_like(q1)
        kv_dev = torch.tensor([keep], dtype=torch.int32, device=dev)
        best = None
        for splits in sorted({8, 16, 37, 74, 111, min(148, (keep + 127) // 128)}):
            ws = torch.zeros(_abi.attn_workspace_bytes(1, HQ, D, splits) // 4 + 16, dtype=torch.float32, device=dev)

            def dec():
                for li in range(L):
                    _abi.attn_causal(c.view, 1, li, q1, o1, 1, HQ, cap, scale, splits, ws, kv_len_dev=kv_dev)
            t = timed(dec) / L
            if best is None or t < best[0]:
                best = (t, splits)
        t, splits = best
        byt = keep * 2 * H * D * 2
        rec["attn_decode"] = {"us_per_layer": t * 1e6, "splits": splits, "GBs": byt / t / 1e9, "frac_hbm": byt / t / 1e9 / pk["hbm_gbs"],
                              "alg_bytes_per_layer": byt}
        if flash_attn_func is not None:
            kf = c.arena[1, 1, 0, :, :keep].transpose(0, 1)[None]      # [1, S, H, D]
            vf = c.arena[1, 2, 0, :, :keep].transpose(0, 1)[None]

            def ref_dec():
                kr = kf[:, :, :, None, :].expand(1, keep, H, HQ // H, D).reshape(1, keep, HQ, D)   # repeat_kv
                vr = vf[:, :, :, None, :].expand(1, keep, H, HQ // H, D).reshape(1, keep, HQ, D)
                flash_attn_func(q1[None], kr, vr, causal=True)
            t = timed(ref_dec, graph=False)
            rec["ref_repeat_kv_plus_flash_attn_decode"] = {"us_per_layer": t * 1e6}

        # chunk prefill: Tq = 276 against n keys
        tq = 276
        qp = torch.randn(tq, HQ, D, device=dev).to(torch.bfloat16)
        op = torch.empty_like(qp)
        fl = 4.0 * tq * (keep - tq / 2) * D * HQ
        best = None
        for splits in (1, 2, 4):
            ws = torch.zeros(_abi.attn_workspace_bytes(tq, HQ, D, splits) // 4 + 16, dtype=torch.float32, device=dev)

            def pre():
                for li in range(L):
                    _abi.attn_causal(c.view, 1, li, qp, op, tq, HQ, keep, scale, splits, ws)
            t = timed(pre) / L
            if best is None or t < best[0]:
                best = (t, splits)
        t, splits = best
        rec["attn_prefill"] = {"us_per_layer": t * 1e6, "splits": splits, "TFLOPs": fl / t / 1e12, "frac_tensor": fl / t / 1e12 / pk["bf16_tflops"],
                               "alg_flop_per_layer": fl}
        if flash_attn_func is not None:
            def ref_pre():
                kr = kf[:, :, :, None, :].expand(1, keep, H, HQ // H, D).reshape(1, keep, HQ, D)
                vr = vf[:, :, :, None, :].expand(1, keep, H, HQ // H, D).reshape(1, keep, HQ, D)
                flash_attn_func(qp[None], kr, vr, causal=True)
            t = timed(ref_pre, graph=False)
            rec["ref_repeat_kv_plus_flash_attn_prefill"] = {"us_per_layer": t * 1e6, "TFLOPs": fl / t / 1e12}
        print(json.dumps(rec), flush=True)
        del c
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
