"""Model-level forwards bound onto an HF Qwen2-VL model (reference `inference/qwen2/model_forward.py`): embed, ViT,
masked_scatter, shrink-mode positions of the FULL pruned ids, language model; then lm_head."""
from __future__ import annotations


def model_forward(self, input_ids=None, past_key_values=None, inputs_embeds=None, pixel_values=None, pixel_values_videos=None,
                  video_grid_thw=None, streaming_args=None, **kwargs):
    from transformers.models.qwen2_vl.modeling_qwen2_vl import Qwen2VLModelOutputWithPast
    if streaming_args is None or streaming_args.pos_mode != "shrink":
        raise RuntimeError("the svlm CUDA path of the patched HF model implements pos_mode='shrink' only")
    if pixel_values is not None:
        raise RuntimeError("images are not part of the streaming path (video chunks only)")
    if inputs_embeds is None:
        inputs_embeds = self.get_input_embeddings()(input_ids)
        if pixel_values_videos is not None:
            feats = self.visual(pixel_values_videos.to(self.visual.patch_embed.proj.weight.dtype), video_grid_thw)
            is_video = input_ids == self.config.video_token_id
            if int(is_video.sum()) != feats.shape[0]:
                raise ValueError(f"Video features and video tokens do not match: tokens: {int(is_video.sum())}, features {feats.shape[0]}")
            inputs_embeds = inputs_embeds.masked_scatter(is_video.unsqueeze(-1).expand_as(inputs_embeds), feats.to(inputs_embeds.device, inputs_embeds.dtype))
    past_key_values.set_video_grids(streaming_args.video_grid_thw.cpu().numpy())
    past_key_values.compute_positions(streaming_args.input_ids)
    out = self.language_model(input_ids=None, past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=True,
                              streaming_args=streaming_args)
    return Qwen2VLModelOutputWithPast(last_hidden_state=out.last_hidden_state, past_key_values=out.past_key_values, rope_deltas=None)


def qwen2_vl_forward(self, input_ids=None, past_key_values=None, inputs_embeds=None, pixel_values_videos=None, video_grid_thw=None,
                     streaming_args=None, logits_to_keep=0, **kwargs):
    from transformers.models.qwen2_vl.modeling_qwen2_vl import Qwen2VLCausalLMOutputWithPast
    out = self.model(input_ids=input_ids, pixel_values_videos=pixel_values_videos, video_grid_thw=video_grid_thw,
                     past_key_values=past_key_values, inputs_embeds=inputs_embeds, streaming_args=streaming_args)
    h = out.last_hidden_state
    if logits_to_keep:
        h = h[:, -logits_to_keep:, :]
    return Qwen2VLCausalLMOutputWithPast(loss=None, logits=self.lm_head(h), past_key_values=out.past_key_values)
