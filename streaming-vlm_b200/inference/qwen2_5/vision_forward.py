"""ViT forwards bound onto an HF Qwen2.5-VL vision tower.

Same three patch points as the reference (`qwen2_5/vision_forward.py`): attention :6-37, block :39-54,
encoder :57-101.  The attention replaces `apply_rotary_pos_emb_flashatt` + `flash_attn_varlen_func` by the
`svlm::vit_attn` op (in-place fp32 RoPE + tcgen05 varlen attention); the merger's two Linears and the
`argsort(window_index)` un-permute run as two `svlm::gemm_bf16` launches with fused bias / GELU / row scatter.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ... import _abi


def streaming_visual_attention_forward(self, hidden_states, cu_seqlens, rotary_pos_emb=None, position_embeddings=None, **kwargs):
    seq_length = hidden_states.shape[0]
    qkv = self.qkv(hidden_states).reshape(seq_length, 3, self.num_heads, -1)
    if position_embeddings is None:
        emb = torch.cat((rotary_pos_emb, rotary_pos_emb), dim=-1)
        cos, sin = emb.cos(), emb.sin()
    else:
        cos, sin = position_embeddings
    out = torch.ops.svlm.vit_attn(qkv, cos.float().contiguous(), sin.float().contiguous(), cu_seqlens.to(torch.int32),
                                  float(qkv.shape[-1]) ** -0.5)
    return self.proj(out.reshape(seq_length, -1))


def streaming_visual_block_forward(self, hidden_states, cu_seqlens, rotary_pos_emb=None, position_embeddings=None, **kwargs):
    hidden_states = hidden_states + self.attn(self.norm1(hidden_states), cu_seqlens=cu_seqlens,
                                              rotary_pos_emb=rotary_pos_emb, position_embeddings=position_embeddings)
    hidden_states = hidden_states + self.mlp(self.norm2(hidden_states))
    return hidden_states


def streaming_visual_encoder_forward(self, hidden_states, grid_thw, **kwargs):
    """patch-embed -> window reorder -> blocks (windowed / full) -> merger -> restore order; returns [N/4, out_hidden]."""
    _abi.register_torch_ops()
    hidden_states = self.patch_embed(hidden_states)
    rotary_pos_emb = self.rot_pos_emb(grid_thw)
    window_index, cu_window_seqlens = self.get_window_index(grid_thw)
    cu_window_seqlens = torch.unique_consecutive(torch.tensor(cu_window_seqlens, device=hidden_states.device, dtype=torch.int32))
    seq_len, _ = hidden_states.size()
    unit = self.spatial_merge_unit
    hidden_states = hidden_states.reshape(seq_len // unit, unit, -1)[window_index].reshape(seq_len, -1)
    rotary_pos_emb = rotary_pos_emb.reshape(seq_len // unit, unit, -1)[window_index].reshape(seq_len, -1)
    emb = torch.cat((rotary_pos_emb, rotary_pos_emb), dim=-1)
    position_embeddings = (emb.cos(), emb.sin())
    cu_seqlens = torch.repeat_interleave(grid_thw[:, 1] * grid_thw[:, 2], grid_thw[:, 0]).cumsum(dim=0).to(torch.int32)
    cu_seqlens = F.pad(cu_seqlens, (1, 0))
    for idx, blk in enumerate(self.blocks):
        cu_now = cu_seqlens if idx in self.fullatt_block_indexes else cu_window_seqlens
        hidden_states = blk(hidden_states, cu_seqlens=cu_now, position_embeddings=position_embeddings)
    m = self.merger
    x = torch.ops.svlm.rmsnorm(hidden_states, m.ln_q.weight, m.ln_q.variance_epsilon).view(-1, m.hidden_size)
    x = torch.ops.svlm.gemm_bf16(x, m.mlp[0].weight, m.mlp[0].bias, 1, None)
    return torch.ops.svlm.gemm_bf16(x, m.mlp[2].weight, m.mlp[2].bias, 0, window_index.to(device=x.device, dtype=torch.int32))
