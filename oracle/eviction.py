"""Oracle: sink / text-window / vision-window KV eviction (test infrastructure, CPU only).

Restates `prune_id_and_kv_cache` (`inference/inference.py:50-61`),
`contiguous_id_and_kv` (`:63-68`) and `process_past_kv` (`:87-172`, including the inner
`resort_id_and_kv`, `:100-108`).

The reference performs up to four full-cache `torch.cat` / `index_select` passes per
chunk.  Every one of those is a gather along the sequence axis, so their composition is
ONE gather.  The oracle therefore tracks a source-index map while it edits the ids and
applies it to the cache once at the end: `new[..., j, :] = old[..., gather[j], :]`.  The
map itself is the bit-exact contract the CUDA compaction kernel is tested against.
"""
from __future__ import annotations

import torch

from .tokens import LF, get_qwen_range


class _Edit:
    """ids plus the map new-index -> index in the cache as it was on entry."""

    def __init__(self, ids):
        self.ids = [int(t) for t in ids]
        self.src = list(range(len(self.ids)))

    def delete(self, a: int, b: int):
        """Remove the closed interval [a, b]  (reference :52, :56)."""
        del self.ids[a:b + 1]
        del self.src[a:b + 1]

    def move_after(self, s: int, e: int, dst: int):
        """Move [s, e] so that it follows index dst  (reference :100-103)."""
        assert dst < s <= e
        for arr in (self.ids, self.src):
            seg = arr[s:e + 1]
            del arr[s:e + 1]
            arr[dst + 1:dst + 1] = seg


def plan_eviction(ids, i, text_round, visual_round, history, assistant_start_bias,
                  assistant_end_bias, clips, pixels, text_sink, text_sliding_window):
    """The id / history side of `process_past_kv`; returns (new_ids list, gather list).

    `ids` is the flat id list of the previous round (prompt + generated, reference
    `prev_generated_ids`).  The cache covers ids[:-1] or ids[:-2] (the last generated
    token(s) were never fed back), which is why the gather is later clipped to the
    cache length -- exactly what `index_select` on a shorter tensor does in the
    reference when the kept indices stay below seq_len (`:55-56` builds the index
    list from the CACHE length, not from the id length).
    """
    ed = _Edit(ids)
    if i >= text_round:
        assert history[0]["role"] == "previous text"
        assert history[-2 * text_round]["role"] == "user"
        assert history[-(2 * text_round - 1)]["role"] == "assistant"
        history[0]["content"] += history[-(2 * text_round - 1)]["content"][:-4]

        a0, a1 = get_qwen_range(ed.ids, "assistant", 0)
        _, p1 = get_qwen_range(ed.ids, "previous text", 0, contain_lf=False)
        s = a0 + assistant_start_bias
        e = a1 - assistant_end_bias - (1 if ed.ids[a1] == LF else 0)
        if s <= e:
            ed.move_after(s, e, p1 - 1)

        content = history[-2 * text_round]["content"]
        for k, item in enumerate(content):
            if item["type"] == "text":
                del content[k]
                break
        del history[-(2 * text_round - 1)]
        if visual_round > text_round:
            u0, u1 = get_qwen_range(ed.ids, "user_text", -text_round, contain_lf=False)
            ed.delete(u0, u1)
        a0, a1 = get_qwen_range(ed.ids, "assistant", -text_round)
        ed.delete(a0, a1)

    if i >= visual_round:
        clips.pop(0)
        pixels.pop(0)
        if visual_round < text_round:
            tgt = history[-2 * visual_round]
            tgt["content"] = [it for it in tgt["content"] if it["type"] != "video"]
            v0, v1 = get_qwen_range(ed.ids, "vision", 0)
            ed.delete(v0, v1)

    if i >= max(visual_round, text_round):
        del history[1]
        u0, u1 = get_qwen_range(ed.ids, "user", 0)
        ed.delete(u0, u1)

    if i > 0 and (text_sink is not None or text_sliding_window is not None):
        p0, p1 = get_qwen_range(ed.ids, "previous text", 0)
        cut0 = p0 + text_sink + 4 if text_sink is not None else p0
        cut1 = p1 - text_sliding_window - 1 if text_sliding_window is not None else p1
        if cut0 <= cut1:
            ed.delete(cut0, cut1)
    return ed.ids, ed.src


def apply_gather(cache, gather):
    """Apply a sequence-axis gather to every layer of a ListCache, in place."""
    if cache is None:
        return cache
    n = cache.get_seq_length()
    idx = torch.tensor([g for g in gather if g < n], dtype=torch.long)
    for li in range(len(cache.key_cache)):
        cache.key_cache[li] = torch.index_select(cache.key_cache[li], 2, idx).contiguous()
        cache.value_cache[li] = torch.index_select(cache.value_cache[li], 2, idx).contiguous()
    return cache


def prune_id_and_kv_cache(input_ids, past_key_values, start_index, end_index):
    """Closed-interval delete from ids [1,S] and every layer's K/V (reference :50-61)."""
    keep = list(range(start_index)) + list(range(end_index + 1, input_ids.shape[1]))
    new_ids = input_ids[:, keep]
    apply_gather(past_key_values, keep)
    return new_ids, past_key_values


def process_past_kv(past_key_values, i, text_round, visual_round, full_conversation_history,
                    prev_generated_ids, assistant_start_bias, assistant_end_bias,
                    recent_video_window_clips, recent_pixel_values_videos, text_sink,
                    text_sliding_window, return_gather=False):
    """Same signature and return tuple as the reference (`inference.py:87-89, 172`)."""
    gather = None
    if prev_generated_ids is not None:
        ids = prev_generated_ids.flatten().tolist()
        new_ids, gather = plan_eviction(
            ids, i, text_round, visual_round, full_conversation_history, assistant_start_bias,
            assistant_end_bias, recent_video_window_clips, recent_pixel_values_videos,
            text_sink, text_sliding_window)
        prev_generated_ids = torch.tensor([new_ids], dtype=prev_generated_ids.dtype)
        apply_gather(past_key_values, gather)
    out = (past_key_values, prev_generated_ids, recent_video_window_clips, recent_pixel_values_videos)
    return out + (gather,) if return_gather else out
