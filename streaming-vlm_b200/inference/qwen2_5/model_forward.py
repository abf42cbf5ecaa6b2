"""Model-level forwards bound onto an HF Qwen2.5-VL model (reference `qwen2_5/model_forward.py`:
`model_forward` :30-176 -- embed, ViT, masked_scatter, shrink-mode positions, language model -- and
`qwen2_5_vl_forward` :221-333 -- lm_head)."""
from __future__ import annotations

import torch


def model_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                  use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None, pixel_values=None,
                  pixel_values_videos=None, image_grid_thw=None, video_grid_thw=None, rope_deltas=None, cache_position=None,
                  second_per_grid_ts=None, streaming_args=None, **kwargs):
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VLModelOutputWithPast
    if streaming_args is None:
        raise RuntimeError("streaming_args is required on the svlm CUDA path")
    if pixel_values is not None:
        raise RuntimeError("images are not part of the streaming path (video chunks only)")
    if inputs_embeds is None:
        inputs_embeds = self.get_input_embeddings()(input_ids)
        if pixel_values_videos is not None:
            video_embeds = self.visual(pixel_values_videos.to(self.visual.patch_embed.proj.weight.dtype), video_grid_thw)
            n_video_tokens = (input_ids == self.config.video_token_id).sum().item()
            if n_video_tokens != video_embeds.shape[0]:
                raise ValueError(f"Video features and video tokens do not match: tokens: {n_video_tokens}, features {video_embeds.shape[0]}")
            mask = (input_ids == self.config.video_token_id).unsqueeze(-1).expand_as(inputs_embeds)
            inputs_embeds = inputs_embeds.masked_scatter(mask, video_embeds.to(inputs_embeds.device, inputs_embeds.dtype))
    cache = past_key_values
    if cache.get_seq_length() == 0:
        cache.pos_mode = streaming_args.pos_mode
    elif cache.pos_mode != streaming_args.pos_mode:
        raise RuntimeError(f"the cache was filled in pos_mode={cache.pos_mode!r}; streaming_args says {streaming_args.pos_mode!r}")
    if streaming_args.pos_mode == "shrink":
        # contiguous positions of the FULL pruned sequence (reference :101-110), written into the cache's device table
        if streaming_args.video_grid_thw is not None:
            cache.set_video_grids(streaming_args.video_grid_thw)
        cache.compute_positions(streaming_args.input_ids, all_text=streaming_args.all_text)
    else:
        # append mode (reference :112-153): positions of the NEW ids only -- get_rope_index over them with the grids of this call,
        # shifted by last_cache_position + 1; a decode step is the one-token case of the same rule.  The running position lives in
        # device memory (`cache.next_pos`); `streaming_args.last_cache_position` is refreshed by `generate` at the end of the turn.
        n_new = inputs_embeds.shape[1]
        cache.reserve_positions(n_new)
        cache.append_new_positions(input_ids, video_grid_thw if pixel_values_videos is not None else None, all_text=streaming_args.all_text)
    outputs = self.language_model(input_ids=None, past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=True,
                                  streaming_args=streaming_args)
    return Qwen2_5_VLModelOutputWithPast(last_hidden_state=outputs.last_hidden_state, past_key_values=outputs.past_key_values,
                                         rope_deltas=None)


def qwen2_5_vl_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                       labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                       pixel_values=None, pixel_values_videos=None, image_grid_thw=None, video_grid_thw=None, rope_deltas=None,
                       cache_position=None, second_per_grid_ts=None, streaming_args=None, logits_to_keep=0, **kwargs):
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VLCausalLMOutputWithPast
    outputs = self.model(input_ids=input_ids, pixel_values_videos=pixel_values_videos, video_grid_thw=video_grid_thw,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, streaming_args=streaming_args)
    hidden_states = outputs.last_hidden_state
    if logits_to_keep:
        hidden_states = hidden_states[:, -logits_to_keep:, :]
    logits = self.lm_head(hidden_states)
    return Qwen2_5_VLCausalLMOutputWithPast(loss=None, logits=logits, past_key_values=outputs.past_key_values)
