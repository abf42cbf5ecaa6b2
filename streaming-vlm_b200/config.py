"""Qwen2.5-VL geometry presets for the hot path.

The reference reads these from the hub checkpoint's `config.json`
(`inference/inference.py:72-77`); offline we carry the public values.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field


@dataclass(frozen=True)
class Geometry:
    name: str
    # language model
    hidden: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    mlp: int
    vocab: int = 152064
    rope_theta: float = 1e6
    mrope_section: tuple = (16, 24, 24)
    rms_eps: float = 1e-6
    # vision tower
    vit_depth: int = 32
    vit_hidden: int = 1280
    vit_heads: int = 16
    vit_mlp: int = 3420
    patch_size: int = 14
    temporal_patch_size: int = 2
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: tuple = (7, 15, 23, 31)
    tokens_per_second: int = 2
    in_channels: int = 3
    vit_kind: str = "qwen2_5"       # "qwen2": the Qwen2-VL tower (LayerNorm, fc1/quick-GELU/fc2, every block full attention)

    @property
    def group(self) -> int:
        return self.num_heads // self.num_kv_heads

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size

    @property
    def kv_bytes_per_token(self) -> int:
        """K + V, all layers, bf16 (57 344 B for 7B)."""
        return self.num_layers * 2 * self.num_kv_heads * self.head_dim * 2

    def as_dict(self) -> dict:
        d = asdict(self)
        d["mrope_section"] = list(self.mrope_section)
        d["fullatt_block_indexes"] = list(self.fullatt_block_indexes)
        return d


TINY = Geometry("tiny", hidden=512, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=128, mlp=1024,
                vit_depth=4, vit_hidden=320, vit_heads=4, vit_mlp=640, fullatt_block_indexes=(1, 3))
QWEN25VL_7B = Geometry("qwen2.5-vl-7b", hidden=3584, num_layers=28, num_heads=28, num_kv_heads=4, head_dim=128,
                       mlp=18944)
# vision tower of the 32B checkpoint: the public 32B / 72B configs carry intermediate_size 3456 (7B: 3420); real checkpoints go
# through `from_hf_config`, which reads the value from the model's own config
QWEN25VL_32B = Geometry("qwen2.5-vl-32b", hidden=5120, num_layers=64, num_heads=40, num_kv_heads=8, head_dim=128,
                        mlp=27648, vit_mlp=3456)

PRESETS = {"tiny": TINY, "7b": QWEN25VL_7B, "32b": QWEN25VL_32B}


def from_hf_config(cfg) -> Geometry:
    """Geometry of a transformers `Qwen2_5_VLConfig` or `Qwen2VLConfig` (HF 5.x field names)."""
    t, v = cfg.text_config, cfg.vision_config
    rp = getattr(t, "rope_parameters", None) or getattr(t, "rope_scaling", None) or {}
    text = dict(hidden=t.hidden_size, num_layers=t.num_hidden_layers, num_heads=t.num_attention_heads,
                num_kv_heads=t.num_key_value_heads, head_dim=t.hidden_size // t.num_attention_heads,
                mlp=t.intermediate_size, vocab=t.vocab_size, rope_theta=float(rp.get("rope_theta", 1e6)),
                mrope_section=tuple(rp.get("mrope_section", (16, 24, 24))), rms_eps=t.rms_norm_eps)
    name = getattr(cfg, "name_or_path", "") or "hf"
    if hasattr(v, "embed_dim"):
        # Qwen2-VL: `embed_dim` is the tower width (`hidden_size` is the merger's output width), no windows, and the temporal
        # position of a vision token is its frame-pair index (no tokens_per_second scaling: qwen2/pos_emb.py)
        return Geometry(name=name, **text, vit_depth=v.depth, vit_hidden=v.embed_dim, vit_heads=v.num_heads,
                        vit_mlp=int(v.embed_dim * v.mlp_ratio), patch_size=v.patch_size, temporal_patch_size=v.temporal_patch_size,
                        spatial_merge_size=v.spatial_merge_size, window_size=0, fullatt_block_indexes=tuple(range(v.depth)),
                        tokens_per_second=1, in_channels=getattr(v, "in_channels", 3), vit_kind="qwen2")
    return Geometry(
        name=name, **text,
        vit_depth=v.depth, vit_hidden=v.hidden_size, vit_heads=v.num_heads, vit_mlp=v.intermediate_size,
        patch_size=v.patch_size, temporal_patch_size=v.temporal_patch_size,
        spatial_merge_size=v.spatial_merge_size, window_size=v.window_size,
        fullatt_block_indexes=tuple(v.fullatt_block_indexes), tokens_per_second=v.tokens_per_second,
        in_channels=getattr(v, "in_channels", 3))
