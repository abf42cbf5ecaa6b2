"""The plugin boundary: bind the B200 path onto a live HF Qwen2.5-VL model.

Same entry point and the same patch points as the reference's `convert_qwen2_5_to_streaming`
(`qwen2_5/patch_model.py:18-38`), written against transformers 5.x module layout (the only version in
this image; the reference pins 4.52.4 -- the kwarg/attribute differences are listed in SURVEY.md 8(b)).
"""
from __future__ import annotations

from types import MethodType

from ... import _abi
from ..generate.streaming_generate_qwen import streaming_generate
from .language_forward import (streaming_language_model_forward, streaming_text_decoder_layer_forward,
                               streaming_text_flash_attn_forward)
from .model_forward import model_forward, qwen2_5_vl_forward
from .pos_emb import get_rope_index
from .vision_forward import (streaming_visual_attention_forward, streaming_visual_block_forward,
                             streaming_visual_encoder_forward)


def convert_qwen2_5_to_streaming(model):
    _abi.lib()                       # fail loudly now if the CUDA library is missing
    _abi.register_torch_ops()
    model.generate = MethodType(streaming_generate, model)
    model.forward = MethodType(qwen2_5_vl_forward, model)
    model.model.forward = MethodType(model_forward, model.model)
    model.model.language_model.forward = MethodType(streaming_language_model_forward, model.model.language_model)
    for layer in model.model.language_model.layers:
        layer.forward = MethodType(streaming_text_decoder_layer_forward, layer)
        layer.self_attn.forward = MethodType(streaming_text_flash_attn_forward, layer.self_attn)
    model.model.visual.forward = MethodType(streaming_visual_encoder_forward, model.model.visual)
    for blk in model.model.visual.blocks:
        blk.forward = MethodType(streaming_visual_block_forward, blk)
        blk.attn.forward = MethodType(streaming_visual_attention_forward, blk.attn)
    model.model.get_rope_index = MethodType(get_rope_index, model.model)
    return model
