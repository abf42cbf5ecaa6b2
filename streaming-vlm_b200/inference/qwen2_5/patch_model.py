"""The plugin boundary: bind the B200 path onto a live HF Qwen2.5-VL model.

Same entry point and the same patch points as the reference's `convert_qwen2_5_to_streaming`
(`qwen2_5/patch_model.py:18-38`), written against transformers 5.x module layout (the only version in
this image; the reference pins 4.52.4 -- the kwarg/attribute differences are listed in SURVEY.md 8(b)).
"""
from __future__ import annotations

from types import MethodType

from ... import _abi
from ..generate import streaming_generate_qwen as _gen
from . import language_forward as _lf
from . import model_forward as _mf
from . import pos_emb as _pe
from . import vision_forward as _vf


def _bind(obj, fn, name: str = "forward"):
    setattr(obj, name, MethodType(fn, obj))


def convert_qwen2_5_to_streaming(model):
    """Rebind generate / forward of the wrapper, the multimodal core, the language model, every decoder layer and its
    attention, the vision tower, every vision block and its attention, and `get_rope_index` -- the reference's patch list."""
    _abi.lib()                       # fail loudly now if the CUDA library is missing
    _abi.register_torch_ops()
    core = model.model
    text, tower = core.language_model, core.visual
    _bind(model, _gen.streaming_generate, "generate")
    _bind(model, _mf.qwen2_5_vl_forward)
    _bind(core, _mf.model_forward)
    _bind(core, _pe.get_rope_index, "get_rope_index")
    _bind(text, _lf.streaming_language_model_forward)
    _bind(tower, _vf.streaming_visual_encoder_forward)
    per_item = ((text.layers, _lf.streaming_text_decoder_layer_forward, "self_attn", _lf.streaming_text_flash_attn_forward),
                (tower.blocks, _vf.streaming_visual_block_forward, "attn", _vf.streaming_visual_attention_forward))
    for modules, outer_fn, inner_name, inner_fn in per_item:
        for m in modules:
            _bind(m, outer_fn)
            _bind(getattr(m, inner_name), inner_fn)
    return model
