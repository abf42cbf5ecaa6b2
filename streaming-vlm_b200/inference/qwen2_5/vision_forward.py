"""ViT forwards bound onto an HF Qwen2.5-VL vision tower.

Same three patch points as the reference (`qwen2_5/vision_forward.py`): attention :6-37, block :39-54,
encoder :57-101.  The attention replaces `apply_rotary_pos_emb_flashatt` + `flash_attn_varlen_func` by the
`svlm::vit_attn` op (in-place fp32 RoPE + tcgen05 varlen attention); the merger's two Linears and the
`argsort(window_index)` un-permute run as two `svlm::gemm_bf16` launches with fused bias / GELU / row scatter.
"""
from __future__ import annotations

import torch

from ... import _abi


def _cos_sin(freqs):
    """HF duplicates the rotary frequencies over both halves of the head."""
    both = torch.cat((freqs, freqs), dim=-1)
    return both.cos(), both.sin()


def streaming_visual_attention_forward(self, hidden_states, cu_seqlens, rotary_pos_emb=None, position_embeddings=None, **kwargs):
    n_tok = hidden_states.shape[0]
    cos, sin = position_embeddings if position_embeddings is not None else _cos_sin(rotary_pos_emb)
    packed = self.qkv(hidden_states).reshape(n_tok, 3, self.num_heads, -1)
    ctx = torch.ops.svlm.vit_attn(packed, cos.float().contiguous(), sin.float().contiguous(), cu_seqlens.to(torch.int32),
                                  float(packed.shape[-1]) ** -0.5)
    return self.proj(ctx.reshape(n_tok, -1))


def streaming_visual_block_forward(self, hidden_states, cu_seqlens, rotary_pos_emb=None, position_embeddings=None, **kwargs):
    attn_in = self.norm1(hidden_states)
    x = hidden_states + self.attn(attn_in, cu_seqlens=cu_seqlens, rotary_pos_emb=rotary_pos_emb, position_embeddings=position_embeddings)
    return x + self.mlp(self.norm2(x))


def _by_window(t, order, unit):
    """Reorder merge units (groups of `unit` consecutive rows) into window-major order."""
    n = t.shape[0]
    return t.reshape(n // unit, unit, -1)[order].reshape(n, -1)


def streaming_visual_encoder_forward(self, hidden_states, grid_thw, **kwargs):
    """patch-embed -> window reorder -> blocks (windowed / full) -> merger -> restore order; returns [N/4, out_hidden]."""
    _abi.register_torch_ops()
    dev = hidden_states.device
    order, win_bounds = self.get_window_index(grid_thw)
    win_bounds = torch.unique_consecutive(torch.tensor(win_bounds, device=dev, dtype=torch.int32))
    per_frame = torch.repeat_interleave(grid_thw[:, 1] * grid_thw[:, 2], grid_thw[:, 0])
    frame_bounds = torch.cat((per_frame.new_zeros(1), per_frame.cumsum(dim=0))).to(torch.int32)
    unit = self.spatial_merge_unit
    x = _by_window(self.patch_embed(hidden_states), order, unit)
    rope = _cos_sin(_by_window(self.rot_pos_emb(grid_thw), order, unit))
    full = set(self.fullatt_block_indexes)
    for li, block in enumerate(self.blocks):
        x = block(x, cu_seqlens=frame_bounds if li in full else win_bounds, position_embeddings=rope)
    head = self.merger
    y = torch.ops.svlm.rmsnorm(x, head.ln_q.weight, head.ln_q.variance_epsilon).view(-1, head.hidden_size)
    y = torch.ops.svlm.gemm_bf16(y, head.mlp[0].weight, head.mlp[0].bias, 1, None)
    # merged row m belongs at row order[m]: the argsort(window_index) un-permute as a row scatter in the GEMM epilogue
    return torch.ops.svlm.gemm_bf16(y, head.mlp[2].weight, head.mlp[2].bias, 0, order.to(device=y.device, dtype=torch.int32))
