"""Shared driver: replays the synthetic stream of tests/golden/make_golden.py::simulate
against any implementation of `process_past_kv` (oracle or product host logic)."""
import hashlib

import numpy as np
import torch

from streaming_vlm_b200 import synthetic


def rle(a):
    runs = []
    for x in a:
        if runs and runs[-1][0] + runs[-1][1] == x:
            runs[-1][1] += 1
        else:
            runs.append([int(x), 1])
    return runs


def sha(ids):
    return hashlib.sha256(np.asarray(ids, dtype=np.int64).tobytes()).hexdigest()[:16]


def replay(process_past_kv, make_cache, read_gather, cfg, recorded_rounds):
    """make_cache(n) -> cache tagged with arange(n); read_gather(cache) -> list of source rows.
    Yields (recorded_round, before_ids, after_ids, gather) for every round with a previous state."""
    hist, clips, pixels = [], [], []
    prev, cache = None, None
    for rec in recorded_rounds:
        i = rec["i"]
        before = None if prev is None else prev.flatten().tolist()
        if cache is not None:
            cache = make_cache(cache_len)
        cache, prev, clips, pixels = process_past_kv(
            cache, i, text_round=cfg["text_round"], visual_round=cfg["window"], full_conversation_history=hist,
            prev_generated_ids=prev, assistant_start_bias=3, assistant_end_bias=2, recent_video_window_clips=clips,
            recent_pixel_values_videos=pixels, text_sink=cfg["text_sink"], text_sliding_window=cfg["text_sliding_window"])
        if before is not None:
            yield rec, before, prev.flatten().tolist(), read_gather(cache)
        clips.append(None)
        prompt = synthetic.round_prompt(i, cfg["n_vis"], cfg["prev_text"], 8, cfg["seed"])
        if i == 0:
            hist[:] = [{"role": "previous text", "content": "p" * 10},
                       {"role": "user", "content": [{"type": "text", "text": "T"}, {"type": "video"}, {"type": "text", "text": "q"}]}]
        else:
            hist.append({"role": "user", "content": [{"type": "text", "text": "T"}, {"type": "video"}]})
        ids = prompt if prev is None else (prev.flatten().tolist() + (prompt if prev[0, -1].item() != 198 else prompt[1:]))
        pixels.append(None)
        assert len(ids) == rec["input_len"]
        seq = ids + rec["gen"]
        cache_len = len(seq) - 1
        if seq[-1] != 151645:
            seq = seq + [151645]
        prev = torch.tensor([seq], dtype=torch.long)
        cache = make_cache(cache_len)
        hist.append({"role": "assistant", "content": "x" * 8 + " ..."})
