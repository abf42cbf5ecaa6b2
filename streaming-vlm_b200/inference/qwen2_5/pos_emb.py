"""Contiguous 3-D mRoPE position ids on the device.

Same call signature and return values as the reference's `get_rope_index`
(`qwen2_5/pos_emb.py:6-13`: `(position_ids float32 [3, B, S], mrope_position_deltas [B, 1])`), for the
case the streaming path uses (`qwen2_5/model_forward.py:101-110`: full pruned ids, no images, all-ones
mask, batch 1).  The reference walks the ids in host Python with `.tolist()` / `.item()` syncs on every
forward; here one single-CTA scan kernel (`svlm_mrope_positions`) produces the integer positions, which
are returned as float32 exactly like the reference's.
"""
from __future__ import annotations

import torch

from ... import _abi


def get_rope_index(self, input_ids=None, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
                   attention_mask=None):
    if input_ids is None or input_ids.shape[0] != 1:
        raise RuntimeError("svlm get_rope_index: batch size 1 only (the streaming path is single-stream)")
    if image_grid_thw is not None:
        raise RuntimeError("svlm get_rope_index: images are not part of the streaming path (video chunks only)")
    if attention_mask is not None and not bool(attention_mask.all()):
        raise RuntimeError("svlm get_rope_index: padding masks are not supported on the streaming path")
    _abi.register_torch_ops()
    cfg = self.config
    ids = input_ids.reshape(-1).to(torch.int64).contiguous()
    n = ids.numel()
    dev = ids.device
    if video_grid_thw is None:
        grids = torch.ones(1, 3, dtype=torch.int32, device=dev)
    else:
        grids = video_grid_thw.to(device=dev, dtype=torch.int32).contiguous()
    pos = torch.empty(3, max(n, 1), dtype=torch.int32, device=dev)
    nxt = torch.empty(1, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    t_step = int(round((2.0 / 2.0) * cfg.vision_config.tokens_per_second))      # second_per_grid_t = 2 / FPS, FPS = 2 (pos_emb.py:104)
    torch.ops.svlm.mrope_positions(ids, grids, pos, nxt, status, cfg.vision_config.spatial_merge_size, t_step, cfg.video_token_id)
    position_ids = pos[:, :n].to(torch.float32).unsqueeze(1)
    deltas = (nxt.to(torch.int64) - n).reshape(1, 1)
    return position_ids, deltas
