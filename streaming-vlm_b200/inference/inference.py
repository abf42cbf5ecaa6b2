"""KV eviction for the streaming session: sink + text window + vision window.

Host-side mirror of the reference's cache functions (`inference/inference.py`):
`prune_id_and_kv_cache` (:50-61), `contiguous_id_and_kv` (:63-68), `process_past_kv`
(:87-172) -- same names, argument meaning, return tuples and assertion behaviour -- but a
different execution model.  The reference edits the cache with up to four full-cache
`torch.cat` / `index_select` passes per chunk, each with a host token scan and a D2H sync.
Here the edits are applied to a host-side (ids, source-row) pair only; their composition is
ONE gather, which is run-length encoded (<= ~8 runs) and executed by a single launch of
`svlm_kv_compact_rope` over all layers, fused with the contiguous-RoPE re-rotation of K.
"""
from __future__ import annotations

import numpy as np
import torch

from ..utils.get_qwen_range import TOKEN_IDS, as_host_ids, find_segment

DEFAULT_WINDOW_SIZE = 16
DEFAULT_TEXT_ROUND = 16
DEFAULT_TEXT_SINK = 512
DEFAULT_TEXT_SLIDING_WINDOW = 512
MAX_TOKEN_PER_DURATION = 20

_LF = TOKEN_IDS["\n"]


class EvictionPlan:
    """ids of the pruned sequence plus, per position, the row it had on entry."""

    def __init__(self, ids):
        self.ids = as_host_ids(ids).copy()
        self.src = np.arange(self.ids.shape[0], dtype=np.int64)

    def _range(self, label, index, contain_lf=True):
        s, e = find_segment(self.ids, label, index, contain_lf)
        return (s, e - 1) if label == "user_text" else (s, e)

    def delete(self, a: int, b: int):
        self.ids = np.concatenate([self.ids[:a], self.ids[b + 1:]])
        self.src = np.concatenate([self.src[:a], self.src[b + 1:]])

    def move_after(self, s: int, e: int, dst: int):
        assert dst < s <= e
        order = lambda x: np.concatenate([x[:dst + 1], x[s:e + 1], x[dst + 1:s], x[e + 1:]])
        self.ids, self.src = order(self.ids), order(self.src)

    def runs(self, cache_len: int):
        """[(src_start, dst_start, length)] over the rows that exist in the cache."""
        src = self.src[self.src < cache_len]
        if src.size == 0:
            return []
        brk = np.flatnonzero(np.diff(src) != 1) + 1
        starts = np.concatenate([[0], brk])
        ends = np.concatenate([brk, [src.size]])
        return [(int(src[a]), int(a), int(b - a)) for a, b in zip(starts, ends)]


def plan_past_kv(ids, i, text_round, visual_round, full_conversation_history, assistant_start_bias, assistant_end_bias,
                 recent_video_window_clips, recent_pixel_values_videos, text_sink, text_sliding_window) -> EvictionPlan:
    """The id / history side of `process_past_kv` (reference :91-170); no device work."""
    plan = EvictionPlan(ids)
    hist = full_conversation_history
    if i >= text_round:
        assert hist[0]["role"] == "previous text"
        assert hist[-2 * text_round]["role"] == "user"
        assert hist[-(2 * text_round - 1)]["role"] == "assistant"
        hist[0]["content"] += hist[-(2 * text_round - 1)]["content"][:-4]
        a0, a1 = plan._range("assistant", 0)
        _, p1 = plan._range("previous text", 0, contain_lf=False)
        s = a0 + assistant_start_bias
        e = a1 - assistant_end_bias - (1 if plan.ids[a1] == _LF else 0)
        if s <= e:
            plan.move_after(s, e, p1 - 1)
        content = hist[-2 * text_round]["content"]
        for k, item in enumerate(content):
            if item["type"] == "text":
                del content[k]
                break
        del hist[-(2 * text_round - 1)]
        if visual_round > text_round:
            plan.delete(*plan._range("user_text", -text_round, contain_lf=False))
        plan.delete(*plan._range("assistant", -text_round))
    if i >= visual_round:
        recent_video_window_clips.pop(0)
        recent_pixel_values_videos.pop(0)
        if visual_round < text_round:
            tgt = hist[-2 * visual_round]
            tgt["content"] = [it for it in tgt["content"] if it["type"] != "video"]
            plan.delete(*plan._range("vision", 0))
    if i >= max(visual_round, text_round):
        del hist[1]
        plan.delete(*plan._range("user", 0))
    if i > 0 and (text_sink is not None or text_sliding_window is not None):
        p0, p1 = plan._range("previous text", 0)
        cut0 = p0 + text_sink + 4 if text_sink is not None else p0
        cut1 = p1 - text_sliding_window - 1 if text_sliding_window is not None else p1
        if cut0 <= cut1:
            plan.delete(cut0, cut1)
    return plan


def _ids_like(ref, ids_np):
    if isinstance(ref, torch.Tensor):
        return torch.from_numpy(ids_np.copy()).to(dtype=ref.dtype).reshape(1, -1).to(ref.device)
    return ids_np.reshape(1, -1)


def prune_id_and_kv_cache(input_ids, past_key_values, start_index, end_index):
    """Delete the closed interval [start_index, end_index] from ids and cache (reference :50-61)."""
    plan = EvictionPlan(input_ids)
    plan.delete(start_index, end_index)
    if past_key_values is not None:
        past_key_values.compact(plan.runs(past_key_values.get_seq_length()), plan.ids)
    return _ids_like(input_ids, plan.ids), past_key_values


def contiguous_id_and_kv(input_ids, past_key_values):
    """The arena is always dense after a compaction; kept for interface parity (reference :63-68)."""
    if isinstance(input_ids, torch.Tensor):
        input_ids = input_ids.contiguous()
    return input_ids, past_key_values


def process_past_kv(past_key_values, i, text_round, visual_round, full_conversation_history, prev_generated_ids,
                    assistant_start_bias, assistant_end_bias, recent_video_window_clips, recent_pixel_values_videos,
                    text_sink, text_sliding_window):
    """Same signature and return tuple as the reference (:87-89, :172)."""
    if prev_generated_ids is None:
        return past_key_values, prev_generated_ids, recent_video_window_clips, recent_pixel_values_videos
    plan = plan_past_kv(prev_generated_ids, i, text_round, visual_round, full_conversation_history, assistant_start_bias,
                        assistant_end_bias, recent_video_window_clips, recent_pixel_values_videos, text_sink,
                        text_sliding_window)
    if past_key_values is not None:
        n = past_key_values.get_seq_length()
        runs = plan.runs(n)
        if not (len(runs) == 1 and runs[0] == (0, 0, n)):        # identity gather: nothing moved
            past_key_values.compact(runs, plan.ids)
    return past_key_values, _ids_like(prev_generated_ids, plan.ids), recent_video_window_clips, recent_pixel_values_videos
