"""Oracle: Qwen2.5-VL vision tower as the reference drives it (test infrastructure, CPU).

Restates `streaming_visual_encoder_forward` (`qwen2_5/vision_forward.py:57-101`),
`streaming_visual_block_forward` (`:39-54`), `streaming_visual_attention_forward`
(`:6-37`) and the HF pieces they call: `rot_pos_emb`, `get_window_index`,
`apply_rotary_pos_emb_flashatt` (fp32 rotate, cast back), `Qwen2_5_VLPatchMerger`
(RMSNorm -> view [N/4, 4*C] -> Linear -> GELU(erf) -> Linear) followed by the
`argsort(window_index)` un-permute (`:98-99`).

Weights are a plain dict with HF parameter names under the prefix `model.visual.`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .attention import varlen_attention


def rmsnorm(x, w, eps=1e-6):
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(x.dtype)


def window_index(grid_thw, window_size=112, spatial_merge_size=2, patch_size=14):
    """HF `get_window_index`: merged-token order grouped by window + cu_seqlens (in ViT
    tokens, before unique_consecutive)."""
    ws = window_size // spatial_merge_size // patch_size
    unit = spatial_merge_size * spatial_merge_size
    order, cu, base = [], [0], 0
    for t, h, w in grid_thw:
        gh, gw = h // spatial_merge_size, w // spatial_merge_size
        idx = torch.arange(t * gh * gw).reshape(t, gh, gw)
        ph, pw = ws - gh % ws, ws - gw % ws
        nh, nw = (gh + ph) // ws, (gw + pw) // ws
        pad = F.pad(idx, (0, pw, 0, ph), value=-100)
        pad = pad.reshape(t, nh, ws, nw, ws).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, ws, ws)
        lens = (pad != -100).sum([2, 3]).reshape(-1)
        flat = pad.reshape(-1)
        order.append(flat[flat != -100] + base)
        cu.extend((lens.cumsum(0) * unit + cu[-1]).tolist())
        base += t * gh * gw
    cu_t = torch.unique_consecutive(torch.tensor(cu, dtype=torch.int32))
    return torch.cat(order), cu_t


def rot_pos_ids(grid_thw, spatial_merge_size=2):
    """HF `rot_pos_emb` index part: (row, col) of every ViT token in merge-unit-major order."""
    out = []
    m = spatial_merge_size
    for t, h, w in grid_thw:
        hp = torch.arange(h)[:, None].expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w)[None, :].expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        out.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
    return torch.cat(out)


def vision_rotary(grid_thw, head_dim, spatial_merge_size=2, theta=10000.0):
    """cos, sin fp32 [N, head_dim] before the window permutation."""
    dim = head_dim // 2
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
    pid = rot_pos_ids(grid_thw, spatial_merge_size)
    mx = max(max(h, w) for _, h, w in grid_thw)
    table = torch.outer(torch.arange(mx, dtype=torch.float), inv)
    emb = table[pid].flatten(1)                  # [N, head_dim/2]
    emb = torch.cat((emb, emb), dim=-1)
    return emb.cos(), emb.sin()


def apply_rope_vision(q, k, cos, sin):
    """q,k [N,H,D]; cos/sin fp32 [N,D]; rotate in fp32, cast back."""
    def rot(x):
        h = x.shape[-1] // 2
        return torch.cat((-x[..., h:], x[..., :h]), dim=-1)
    c, s = cos[:, None, :].float(), sin[:, None, :].float()
    qf, kf = q.float(), k.float()
    return ((qf * c) + (rot(qf) * s)).to(q.dtype), ((kf * c) + (rot(kf) * s)).to(k.dtype)


def vit_attention(x, w, pfx, num_heads, cu, cos, sin):
    n = x.shape[0]
    qkv = F.linear(x, w[pfx + "qkv.weight"], w[pfx + "qkv.bias"]).reshape(n, 3, num_heads, -1)
    q, k, v = qkv.permute(1, 0, 2, 3).unbind(0)
    q, k = apply_rope_vision(q, k, cos, sin)
    o = varlen_attention(q, k, v, cu).reshape(n, -1)
    return F.linear(o, w[pfx + "proj.weight"], w[pfx + "proj.bias"])


def vit_mlp(x, w, pfx):
    g = F.linear(x, w[pfx + "gate_proj.weight"], w[pfx + "gate_proj.bias"])
    u = F.linear(x, w[pfx + "up_proj.weight"], w[pfx + "up_proj.bias"])
    return F.linear(F.silu(g) * u, w[pfx + "down_proj.weight"], w[pfx + "down_proj.bias"])


def patch_merger(x, w, pfx="model.visual.merger.", unit=4):
    h = rmsnorm(x, w[pfx + "ln_q.weight"], 1e-6)
    h = h.view(-1, h.shape[-1] * unit)
    h = F.linear(h, w[pfx + "mlp.0.weight"], w[pfx + "mlp.0.bias"])
    h = F.gelu(h)
    return F.linear(h, w[pfx + "mlp.2.weight"], w[pfx + "mlp.2.bias"])


def vision_forward_qwen2(pixels, grid_thw, w, geo):
    """The Qwen2-VL tower the reference's `inference/qwen2/vision_forward.py:57-80` drives (HF
    `Qwen2VisionTransformerPretrainedModel`): patch embed -> 2-D rotary -> blocks of LayerNorm + full attention per frame +
    LayerNorm + fc1 / quick-GELU / fc2 -> `PatchMerger` (LayerNorm -> view [N/4, 4C] -> Linear -> GELU(erf) -> Linear).
    No window reordering.  Pinned against the HF module itself in tests/test_oracle_golden.py."""
    pfx = "model.visual."
    unit = geo["spatial_merge_size"] ** 2
    wt = w[pfx + "patch_embed.proj.weight"]
    x = F.linear(pixels.to(wt.dtype), wt.reshape(wt.shape[0], -1))
    head_dim = geo["vit_hidden"] // geo["vit_heads"]
    cos, sin = vision_rotary(grid_thw, head_dim, geo["spatial_merge_size"])
    full = [0]
    for t, h, ww in grid_thw:
        for _ in range(t):
            full.append(full[-1] + h * ww)
    cu = torch.tensor(full, dtype=torch.int32)
    ln = lambda t_, name: F.layer_norm(t_, (t_.shape[-1],), w[name + ".weight"], w[name + ".bias"], 1e-6)
    for li in range(geo["vit_depth"]):
        b = f"{pfx}blocks.{li}."
        x = x + vit_attention(ln(x, b + "norm1"), w, b + "attn.", geo["vit_heads"], cu, cos, sin)
        h1 = F.linear(ln(x, b + "norm2"), w[b + "mlp.fc1.weight"], w[b + "mlp.fc1.bias"])
        x = x + F.linear(h1 * torch.sigmoid(1.702 * h1), w[b + "mlp.fc2.weight"], w[b + "mlp.fc2.bias"])      # QuickGELU
    m = pfx + "merger."
    y = ln(x, m + "ln_q").view(-1, x.shape[-1] * unit)
    y = F.gelu(F.linear(y, w[m + "mlp.0.weight"], w[m + "mlp.0.bias"]))
    return F.linear(y, w[m + "mlp.2.weight"], w[m + "mlp.2.bias"])


def vision_forward(pixels, grid_thw, w, geo, return_hidden=False, timers=None):
    """pixels [N, C*Tp*P*P], grid_thw list of (t,h,w) -> merged tokens [N/4, out_hidden]."""
    if geo.get("vit_kind", "qwen2_5") == "qwen2":
        return vision_forward_qwen2(pixels, grid_thw, w, geo)
    pfx = "model.visual."
    unit = geo["spatial_merge_size"] ** 2
    wt = w[pfx + "patch_embed.proj.weight"]
    x = F.linear(pixels.to(wt.dtype), wt.reshape(wt.shape[0], -1))
    n = x.shape[0]
    head_dim = geo["vit_hidden"] // geo["vit_heads"]
    cos, sin = vision_rotary(grid_thw, head_dim, geo["spatial_merge_size"])
    widx, cu_win = window_index(grid_thw, geo["window_size"], geo["spatial_merge_size"], geo["patch_size"])
    x = x.reshape(n // unit, unit, -1)[widx].reshape(n, -1)
    cos = cos.reshape(n // unit, unit, -1)[widx].reshape(n, -1)
    sin = sin.reshape(n // unit, unit, -1)[widx].reshape(n, -1)
    full = [0]
    for t, h, ww in grid_thw:
        for _ in range(t):
            full.append(full[-1] + h * ww)
    cu_full = torch.tensor(full, dtype=torch.int32)
    hidden = []
    import time as _time
    _t0 = _time.perf_counter()
    for li in range(geo["vit_depth"]):
        b = f"{pfx}blocks.{li}."
        cu = cu_full if li in geo["fullatt_block_indexes"] else cu_win
        x = x + vit_attention(rmsnorm(x, w[b + "norm1.weight"]), w, b + "attn.", geo["vit_heads"], cu, cos, sin)
        x = x + vit_mlp(rmsnorm(x, w[b + "norm2.weight"]), w, b + "mlp.")
        if return_hidden:
            hidden.append(x)
    if timers is not None:
        timers["vit_blocks"] = timers.get("vit_blocks", 0.0) + _time.perf_counter() - _t0
    merged = patch_merger(x, w, pfx + "merger.", unit)
    merged = merged[torch.argsort(widx)]
    return (merged, hidden) if return_hidden else merged
