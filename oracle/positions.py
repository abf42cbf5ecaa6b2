"""Oracle: contiguous 3-D mRoPE positions and rotation (test infrastructure, CPU only).

Restates
  * `get_rope_index` (`qwen2_5/pos_emb.py:6-188`) as the shrink-mode caller uses it
    (`qwen2_5/model_forward.py:101-110`: full pruned ids, no images, all-ones mask),
  * HF `Qwen2_5_VLRotaryEmbedding.forward` (called at `qwen2_5/language_forward.py:454`):
    fp32 `inv_freq * pos`, duplicated halves, cos/sin, cast to the model dtype,
  * `apply_multimodal_rotary_pos_emb` (`qwen2_5/language_forward.py:11-67`): section
    select [16,24,24]*2 over the (t,h,w) axes, `x*cos + rotate_half(x)*sin` with every
    product and the sum rounded to the tensor dtype (bf16), q right-aligned (`:47-56`).
"""
from __future__ import annotations

import numpy as np
import torch

from .tokens import VIDEO_PAD, VISION_START

FPS = 2.0  # qwen_vl_utils.vision_process.FPS (pinned qwen-vl-utils 0.0.11)


def get_rope_index(ids, video_grid_thw, spatial_merge_size=2, tokens_per_second=2,
                   video_token_id=VIDEO_PAD, vision_start_token_id=VISION_START):
    """ids: flat list/array [S]; video_grid_thw: [[t,h,w], ...] consumed from index 0.

    Returns (position_ids float32 [3, 1, S], delta int).  Follows pos_emb.py:72-158.
    """
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    n = ids.shape[0]
    grids = np.asarray(video_grid_thw, dtype=np.int64).reshape(-1, 3) if video_grid_thw is not None else np.zeros((0, 3), np.int64)
    starts = np.nonzero(ids == vision_start_token_id)[0]
    starts = starts[starts + 1 < n] if starts.size else starts
    n_videos = int((ids[starts + 1] == video_token_id).sum()) if starts.size else 0

    pieces = []
    st = 0
    nxt = 0                      # next free position = previous max + 1   (:117, :140)
    sec_per_grid_t = np.float32(2.0 / FPS)                                 # (:104)
    for vi in range(n_videos):
        where = np.nonzero(ids[st:] == video_token_id)[0]
        ed = st + int(where[0])                                            # (:92, :108)
        t, h, w = (int(x) for x in grids[vi])
        gh, gw = h // spatial_merge_size, w // spatial_merge_size          # (:111-113)
        text_len = ed - st
        text = np.arange(text_len, dtype=np.float32) + nxt
        pieces.append(np.stack([text, text, text]))                        # (:118-120)
        base = np.float32(text_len + nxt)
        tt = (np.arange(t, dtype=np.float32)[:, None] * sec_per_grid_t * np.float32(tokens_per_second))
        tt = np.broadcast_to(tt, (t, gh * gw)).reshape(-1)                 # (:123-127)
        hh = np.broadcast_to(np.arange(gh, dtype=np.float32)[None, :, None], (t, gh, gw)).reshape(-1)
        ww = np.broadcast_to(np.arange(gw, dtype=np.float32)[None, None, :], (t, gh, gw)).reshape(-1)
        block = np.stack([tt, hh, ww]) + base                              # (:129-134)
        pieces.append(block)
        nxt = float(block.max()) + 1 if block.size else (float(text.max()) + 1 if text.size else nxt)
        st = ed + t * gh * gw                                              # (:136)
    if st < n:                                                             # (:139-144)
        text = np.arange(n - st, dtype=np.float32) + np.float32(nxt)
        pieces.append(np.stack([text, text, text]))
    if pieces:
        pos = np.concatenate(pieces, axis=1).astype(np.float32)
    else:
        pos = np.zeros((3, 0), np.float32)
    delta = int(pos.max() + 1 - n) if pos.size else 0                      # (:154)
    return torch.from_numpy(np.ascontiguousarray(pos)).reshape(3, 1, n), delta


def default_inv_freq(head_dim=128, theta=1e6):
    """HF `compute_default_rope_parameters`: 1 / theta^(2i/d), fp32."""
    return 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float32) / head_dim))


def rotary_cos_sin(position_ids, inv_freq, dtype=torch.bfloat16):
    """position_ids float [3,1,S] -> cos, sin [3,1,S,D] in `dtype` (HF rotary forward)."""
    inv = inv_freq[None, None, :, None].float().expand(3, position_ids.shape[1], -1, 1)
    pos = position_ids[:, :, None, :].float()
    freqs = (inv @ pos).transpose(2, 3)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotary_table(max_pos, inv_freq, dtype=torch.bfloat16):
    """cos/sin [max_pos, D/2] for integer positions 0..max_pos-1, same arithmetic as
    `rotary_cos_sin` (fp32 product of inv_freq and the position, fp32 cos/sin, one
    rounding to `dtype`).  The two halves of HF's `emb` are identical, so D/2 suffices."""
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :].float()
    return freqs.cos().to(dtype), freqs.sin().to(dtype)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def select_sections(t, mrope_section):
    """[3,1,S,D] -> [1,1,S,D]: chunk j of `mrope_section*2` comes from axis j % 3."""
    parts = t.split(list(mrope_section) * 2, dim=-1)
    return torch.cat([m[j % 3] for j, m in enumerate(parts)], dim=-1).unsqueeze(1)


def apply_mrope(q, k, cos, sin, mrope_section=(16, 24, 24)):
    """q [1,Hq,Tq,D], k [1,Hkv,S,D], cos/sin [3,1,S,D]; q takes the last Tq positions."""
    c = select_sections(cos, mrope_section)
    s = select_sections(sin, mrope_section)
    tq = q.shape[2]
    qc, qs = c[:, :, -tq:], s[:, :, -tq:]
    q_rot = (q * qc) + (rotate_half(q) * qs)
    k_rot = (k * c) + (rotate_half(k) * s)
    return q_rot, k_rot


def append_mode_positions(new_ids, chunk_grids, cache_len: int, state: dict, spatial_merge_size=2, tokens_per_second=2):
    """`pos_mode="append"` (qwen2_5/model_forward.py:112-153): positions of the NEW tokens only.  Empty cache: plain
    `get_rope_index` over the ids (:114-127); chunk prefill: `get_rope_index` over the new ids with this chunk's grids, shifted by
    `last_cache_position + 1` (:128-137); single token: `last_cache_position + 1` on all three axes (:138-151).  Updates
    `state["last_cache_position"]` with the time-axis position of the last token (:153).  Returns float32 [3, 1, T]."""
    if cache_len == 0:
        pos, _ = get_rope_index(new_ids, chunk_grids, spatial_merge_size, tokens_per_second)
    elif len(new_ids) != 1:
        pos, _ = get_rope_index(new_ids, chunk_grids, spatial_merge_size, tokens_per_second)
        pos = pos + float(state["last_cache_position"] + 1)
    else:
        pos = torch.full((3, 1, 1), float(state["last_cache_position"] + 1))
    state["last_cache_position"] = int(pos[0, 0, -1])
    return pos
