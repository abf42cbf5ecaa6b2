"""LLM forwards bound onto an HF Qwen2.5-VL text model (both position modes of `StreamingArgs`).

Patch points of the reference (`qwen2_5/language_forward.py`): attention :143-260, decoder layer :262-305,
language model :395-516.  The attention keeps the reference's order -- project, cache the RAW keys, rotate
at the contiguous positions of the pruned sequence, causal bottom-right attention over the whole cache,
o_proj -- but appends in place into the `StreamingCache` arena and rotates only the NEW rows
(`svlm::qkv_rope_append`); the cached rows were rotated when they were appended or last compacted.
`repeat_kv` and the per-forward whole-cache rotation do not exist here.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ... import _abi
from ..generate.streaming_cache import StreamingCache


def streaming_text_flash_attn_forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                                      output_attentions=False, use_cache=False, cache_position=None, position_embeddings=None,
                                      streaming_args=None, past_key_values=None, **kwargs):
    cache = past_key_value if past_key_value is not None else past_key_values
    if output_attentions:
        raise RuntimeError("output_attentions=True is not available on the svlm CUDA path (no eager fallback)")
    if streaming_args is None:
        raise RuntimeError("streaming_args is required on the svlm CUDA path")
    if not isinstance(cache, StreamingCache):
        raise RuntimeError("past_key_values must be a streaming_vlm_b200 StreamingCache")
    bsz, q_len, _ = hidden_states.size()
    if bsz != 1:
        raise RuntimeError("batch size 1 only")
    hq, d = self.config.num_attention_heads, self.head_dim
    x = hidden_states[0]
    q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
    start = cache.layer_start(self.layer_idx, q_len)
    q_rot = torch.empty(q_len, hq, d, dtype=q.dtype, device=q.device)
    torch.ops.svlm.qkv_rope_append(cache.arena, cache.cur, self.layer_idx, q, k, v, q_rot, start, cache.pos, cache.cos_table,
                                   cache.sin_table, list(cache.section))
    splits = 1 if q_len > 1 else cache.decode_splits
    out = torch.ops.svlm.attn_causal(cache.arena, cache.cur, self.layer_idx, q_rot, start + q_len, float(d) ** -0.5, splits)
    attn_output = self.o_proj(out.reshape(1, q_len, -1))
    return attn_output, None, cache


def streaming_text_decoder_layer_forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                                         output_attentions=False, use_cache=False, cache_position=None, position_embeddings=None,
                                         streaming_args=None, past_key_values=None, **kwargs):
    cache = past_key_value if past_key_value is not None else past_key_values
    residual = hidden_states
    hidden_states = self.input_layernorm(hidden_states)
    hidden_states, _, _ = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                                         past_key_value=cache, use_cache=use_cache, cache_position=cache_position,
                                         position_embeddings=position_embeddings, streaming_args=streaming_args)
    hidden_states = residual + hidden_states
    residual = hidden_states
    hidden_states = self.mlp(self.post_attention_layernorm(hidden_states))
    hidden_states = residual + hidden_states
    outputs = (hidden_states,)
    if use_cache:
        outputs += (cache,)
    return outputs


def streaming_language_model_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                                     inputs_embeds=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                                     return_dict=None, cache_position=None, streaming_args=None, **kwargs):
    from transformers.modeling_outputs import BaseModelOutputWithPast
    if (input_ids is None) ^ (inputs_embeds is not None):
        raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
    if past_key_values is None:
        raise RuntimeError("pass a streaming_vlm_b200 StreamingCache as past_key_values (it owns the HBM arena)")
    if inputs_embeds is None:
        inputs_embeds = self.embed_tokens(input_ids)
    hidden_states = inputs_embeds
    for decoder_layer in self.layers:
        hidden_states = decoder_layer(hidden_states, past_key_value=past_key_values, use_cache=True, streaming_args=streaming_args)[0]
    hidden_states = self.norm(hidden_states)
    if streaming_args.pos_mode == "shrink":
        # one pad per forward keeps `streaming_args.input_ids` as long as the cache (reference :505-507)
        streaming_args.input_ids = F.pad(streaming_args.input_ids, (0, 1), "constant", 0)
    return BaseModelOutputWithPast(last_hidden_state=hidden_states, past_key_values=past_key_values)
