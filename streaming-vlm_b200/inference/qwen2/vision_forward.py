"""ViT forwards bound onto an HF Qwen2-VL vision tower (reference `inference/qwen2/vision_forward.py`: attention :6-37,
block :39-54, encoder :57-80).  The Qwen2-VL tower has no windows: every block attends within a frame, so the attention and
block forwards of the Qwen2.5 variant apply unchanged (`svlm::vit_attn`: in-place fp32 RoPE + tcgen05 varlen attention with
`head_dim = 80` zero-padded by TMA); LayerNorm, the fc1 / quick-GELU / fc2 MLP and the `PatchMerger` stay HF modules."""
from __future__ import annotations

import torch

from ... import _abi
from ..qwen2_5.vision_forward import (_cos_sin, streaming_visual_attention_forward,  # noqa: F401  (re-exported patch points)
                                      streaming_visual_block_forward)


def streaming_visual_encoder_forward(self, hidden_states, grid_thw, **kwargs):
    """patch-embed -> blocks (full attention per frame) -> merger; returns the merged tokens [N/4, out_hidden]."""
    _abi.register_torch_ops()
    x = self.patch_embed(hidden_states)
    rope = _cos_sin(self.rot_pos_emb(grid_thw))
    per_frame = torch.repeat_interleave(grid_thw[:, 1] * grid_thw[:, 2], grid_thw[:, 0])
    frame_bounds = torch.cat((per_frame.new_zeros(1), per_frame.cumsum(dim=0))).to(torch.int32)
    for block in self.blocks:
        x = block(x, cu_seqlens=frame_bounds, position_embeddings=rope)
    return self.merger(x)
