"""The plugin boundary for Qwen2-VL: `convert_qwen2_to_streaming` (reference `inference/qwen2/patch_model.py`), same patch
points as the Qwen2.5 variant.  The language model is the same Qwen2 decoder with mRoPE, so its forwards are shared."""
from __future__ import annotations

from ... import _abi
from ..generate import streaming_generate_qwen as _gen
from ..qwen2_5 import language_forward as _lf
from ..qwen2_5.patch_model import _bind
from . import model_forward as _mf
from . import pos_emb as _pe
from . import vision_forward as _vf


def convert_qwen2_to_streaming(model):
    _abi.lib()                       # fail loudly now if the CUDA library is missing
    _abi.register_torch_ops()
    core = model.model
    text, tower = core.language_model, core.visual
    _bind(model, _gen.streaming_generate, "generate")
    _bind(model, _mf.qwen2_vl_forward)
    _bind(core, _mf.model_forward)
    _bind(core, _pe.get_rope_index, "get_rope_index")
    _bind(text, _lf.streaming_language_model_forward)
    _bind(tower, _vf.streaming_visual_encoder_forward)
    for layer in text.layers:
        _bind(layer, _lf.streaming_text_decoder_layer_forward)
        _bind(layer.self_attn, _lf.streaming_text_flash_attn_forward)
    for block in tower.blocks:
        _bind(block, _vf.streaming_visual_block_forward)
        _bind(block.attn, _vf.streaming_visual_attention_forward)
    return model
