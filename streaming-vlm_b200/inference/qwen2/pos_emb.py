"""Contiguous 3-D mRoPE position ids on the device, Qwen2-VL flavour (reference `inference/qwen2/pos_emb.py`): identical to
the Qwen2.5 variant except that the temporal index of a vision token is its frame-pair index (no `tokens_per_second` /
`second_per_grid_ts` scaling), i.e. a temporal step of 1."""
from __future__ import annotations

import torch

from ... import _abi


def get_rope_index(self, input_ids=None, image_grid_thw=None, video_grid_thw=None, attention_mask=None, **kwargs):
    if input_ids is None or input_ids.shape[0] != 1:
        raise RuntimeError("svlm get_rope_index: batch size 1 only (the streaming path is single-stream)")
    if image_grid_thw is not None:
        raise RuntimeError("svlm get_rope_index: images are not part of the streaming path (video chunks only)")
    if attention_mask is not None and not bool(attention_mask.all()):
        raise RuntimeError("svlm get_rope_index: padding masks are not supported on the streaming path")
    _abi.register_torch_ops()
    cfg = self.config
    ids = input_ids.reshape(-1).to(torch.int64).contiguous()
    n, dev = ids.numel(), ids.device
    grids = torch.ones(1, 3, dtype=torch.int32, device=dev) if video_grid_thw is None else video_grid_thw.to(device=dev, dtype=torch.int32).contiguous()
    pos = torch.empty(3, max(n, 1), dtype=torch.int32, device=dev)
    nxt = torch.empty(1, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    torch.ops.svlm.mrope_positions(ids, grids, pos, nxt, status, cfg.vision_config.spatial_merge_size, 1, cfg.video_token_id)
    return pos[:, :n].to(torch.float32).unsqueeze(1), (nxt.to(torch.int64) - n).reshape(1, 1)
