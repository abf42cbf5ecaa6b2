#!/usr/bin/env python
"""One-hour synthetic stream (3600 one-second chunks at 2 fps) of the 7B architecture THROUGH THE DROP-IN API -- the reference's loop
body (`process_past_kv` + patched `model.generate`, bench.DropInSession) with the reference's generation config -- from an EMPTY
cache, with turns of varying length: the token budget of a turn cycles through 1..20 and the prompt's time string through 1..9
tokens, so the new-token count of every chunk prefill and the retained length keep changing.  Per video-minute: p50 / p99 chunk
latency, frames/s, retained cache length, number of captured CUDA graphs (bounded LRU caches) and allocated HBM.

    python benchmarks/soak_dropin.py [chunks] > profiles/r2e_soak_dropin_1h_7b.jsonl
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from streaming_vlm_b200 import _abi, synthetic  # noqa: E402
from streaming_vlm_b200.config import QWEN25VL_7B as G7  # noqa: E402
from streaming_vlm_b200.inference.generate.streaming_generate_qwen import _fast_path_engine  # noqa: E402
from streaming_vlm_b200.inference.qwen2_5.patch_model import convert_qwen2_5_to_streaming  # noqa: E402


def main():
    chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
    fixed = len(sys.argv) > 2 and sys.argv[2] == "fixed"          # fixed: 20-token turns, constant prompt length (the bench workload)
    _abi.build()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    w = synthetic.random_weights(G7, seed=0, device=dev)
    model = convert_qwen2_5_to_streaming(bench.build_hf_model(G7, w, dev))
    del w
    engine = _fast_path_engine(model, None, 8192)
    sup = synthetic.suppress_mask(G7.vocab).to(dev)
    engine.set_suppress(sup)
    state = {"prev_ids": None, "history": [], "clips": [], "pixels": [], "grids": [], "i": 0}
    sess = bench.DropInSession.__new__(bench.DropInSession)
    from streaming_vlm_b200.inference.streaming_args import StreamingArgs
    sess.model, sess.geo, sess.dev, sess.sup = model, G7, dev, sup
    sess.gen_kwargs = dict(do_sample=True, temperature=0.9, repetition_penalty=1.05)
    sess.args = StreamingArgs(pos_mode="shrink")
    sess.past, sess.prev, sess.history, sess.clips, sess.pixels, sess.i, sess.timing = None, None, [], [], [], 0, []
    sess.args.video_grid_thw = torch.zeros(0, 3, dtype=torch.long, device=dev)
    frames = [synthetic.frames(k).permute(0, 2, 3, 1).contiguous().pin_memory() for k in range(64)]
    lat, t_wall0 = [], time.perf_counter()
    clk = bench.ClockSampler(0)
    clk.__enter__()
    for i in range(chunks):
        budget = 20 if fixed else 1 + (i * 7) % 20
        prompt = synthetic.round_prompt(i, time_tokens=8 if fixed else 1 + (i * 5) % 9)
        t0 = time.perf_counter()
        if i == 0:      # the first chunk of the reference's loop: no previous ids, history created here (inference.py:351-372)
            first_step(sess, prompt, frames[0].to(dev, non_blocking=True), budget)
        else:
            sess.step(prompt, frames[i % 64].to(dev, non_blocking=True), max_new_tokens=budget)
        torch.cuda.current_stream().synchronize()
        lat.append(time.perf_counter() - t0)
        if (i + 1) % 60 == 0:
            window = lat[-60:]
            print(json.dumps({"video_minute": (i + 1) // 60, "p50_chunk_ms": 1e3 * statistics.median(window), "p99_chunk_ms": 1e3 * sorted(window)[-1],
                              "frames_per_s": 120.0 / sum(window), "kv_len": sess.past.get_seq_length(), "new_tokens_in_minute": sum(t["new_tokens"] for t in sess.timing[-60:]),
                              "prefill_graphs": len(engine._pf_graphs), "decode_graphs": len(engine._graphs), "graph_evictions": engine._pf_graphs.evictions + engine._graphs.evictions,
                              "hbm_allocated_GB": torch.cuda.memory_allocated() / 1e9, "hbm_reserved_GB": torch.cuda.memory_reserved() / 1e9,
                              "compactions": sess.past.compactions, "wall_s": time.perf_counter() - t_wall0}), flush=True)
    clk.__exit__()
    print(json.dumps({"summary": True, "chunks": chunks, "wall_s": time.perf_counter() - t_wall0, "p50_chunk_ms_all": 1e3 * statistics.median(lat),
                      "p50_first_tenth_after_fill_ms": 1e3 * statistics.median(lat[chunks // 10:2 * chunks // 10]) if chunks >= 20 else None,
                      "p50_last_tenth_ms": 1e3 * statistics.median(lat[-max(chunks // 10, 1):]), "clocks": clk.summary()}), flush=True)


def first_step(sess, prompt, frames_thwc, budget):
    """Chunk 0 of the reference's loop: nothing cached, the conversation history starts here."""
    sess.history[:] = [{"role": "previous text", "content": ""},
                       {"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}, {"type": "text", "text": "query"}]}]
    sess.clips.append(None)
    sess.pixels.append(None)
    px = torch.ops.svlm.frames_to_patches(frames_thwc, G7.patch_size, G7.temporal_patch_size, G7.spatial_merge_size)
    grid = torch.tensor([[1, frames_thwc.shape[1] // G7.patch_size, frames_thwc.shape[2] // G7.patch_size]], dtype=torch.long)
    ids = torch.tensor([prompt], dtype=torch.long, device=sess.dev)
    sess.args.input_ids = ids
    sess.args.video_grid_thw = grid.to(sess.dev)
    out = sess.model.generate(input_ids=ids, pixel_values_videos=px, video_grid_thw=grid, past_key_values=None, max_new_tokens=budget, use_cache=True,
                              return_dict_in_generate=True, streaming_args=sess.args, pad_token_id=151645, suppress_tokens=sess.sup, **sess.gen_kwargs)
    seq = out.sequences
    if seq[0, -1].item() != 151645:
        seq = torch.cat([seq, torch.tensor([[151645]], device=sess.dev)], dim=1)
    sess.past, sess.prev = out.past_key_values, seq.clone()
    new = seq.shape[1] - ids.shape[1]
    sess.history.append({"role": "assistant", "content": "x " * new + "..."})
    sess.timing.append({"PKV": 0.0, "GEN": 0.0, "new_tokens": new})
    sess.i = 1


if __name__ == "__main__":
    main()
