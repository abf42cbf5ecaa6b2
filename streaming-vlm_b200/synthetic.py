"""Synthetic streams: token grammar, frames and random-init weights.

No tokenizer, processor, checkpoint or video exists offline, so the benchmark and the
parity tests drive the path with generated inputs of the reference's exact shape
(SURVEY.md section 8(d)):

* token grammar of one streaming round as `streaming_inference` builds it
  (`inference/inference.py:351-406`): system prompt, the "previous text" block, then per
  round `<|im_start|>user\\nTime=...<|vision_start|><|video_pad|>*n<|vision_end|>[query]
  <|im_end|>\\n<|im_start|>assistant\\n`, rounds > 0 prefixed by "\\n" (`:388`);
* frames: uint8 `[2, 3, H, W]` per 1-second chunk at 2 fps, normalised and patchified in
  the HF Qwen2-VL image-processor order (merge-unit-major) into
  `pixel_values_videos [grid_h*grid_w, 3*2*14*14]`, `video_grid_thw = [[1, H/14, W/14]]`;
* weights: normal(0, std) tensors under HF parameter names.
"""
from __future__ import annotations

import math

import torch

from .config import Geometry
from .utils.get_qwen_range import TOKEN_IDS

IM_START, IM_END = TOKEN_IDS["<|im_start|>"], TOKEN_IDS["<|im_end|>"]
LF = TOKEN_IDS["\n"]
SYSTEM_PROMPT = [IM_START, 8948, LF, 2610, 525, 264, 10950, 17847, 13, IM_END, LF]  # 11 tokens
QUERY = [1092, 78182, 389, 419, 2432]                                                 # "Commentate on this match"
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _text_tokens(n: int, seed: int):
    """n plain-text ids in [1000, 100000) that never collide with structural tokens."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(1000, 100000, (n,), generator=g).tolist()
    bad = {TOKEN_IDS["user"], TOKEN_IDS["assistant"], TOKEN_IDS["Time"], LF, *TOKEN_IDS["previous text"]}
    return [x if x not in bad else 1000 for x in t]


def round_prompt(i: int, n_vision_tokens: int = 256, previous_text_tokens: int = 24, time_tokens: int = 8,
                 seed: int = 0):
    """Token ids the processor would return for round i (before concatenation with the
    previous ids).  Round 0 carries system + previous-text + query."""
    user = [IM_START, TOKEN_IDS["user"], LF, TOKEN_IDS["Time"]] + _text_tokens(time_tokens, seed * 1000003 + i)
    user += [TOKEN_IDS["<|vision_start|>"]] + [TOKEN_IDS["<|video_pad|>"]] * n_vision_tokens + [TOKEN_IDS["<|vision_end|>"]]
    tail = [IM_END, LF, IM_START, TOKEN_IDS["assistant"], LF]
    if i == 0:
        prev = [IM_START] + TOKEN_IDS["previous text"] + [LF] + _text_tokens(previous_text_tokens, seed + 77) + [IM_END, LF]
        return SYSTEM_PROMPT + prev + user + QUERY + tail
    return [LF] + user + tail


def suppress_mask(vocab: int, device="cpu"):
    """Ids a random-init model must never emit so that the grammar stays parseable: every
    special token >= 151643 except <|im_end|>, plus the structural text ids."""
    m = torch.zeros(vocab, dtype=torch.bool, device=device)
    m[151643:] = True
    m[IM_END] = False
    for t in (TOKEN_IDS["Time"], TOKEN_IDS["user"], TOKEN_IDS["assistant"], LF, *TOKEN_IDS["previous text"]):
        m[t] = True
    return m


def frames(i: int, height: int = 448, width: int = 448, n: int = 2, seed: int = 1234):
    """uint8 [n, 3, H, W] for chunk i."""
    g = torch.Generator().manual_seed(seed + i)
    return torch.randint(0, 256, (n, 3, height, width), generator=g, dtype=torch.uint8)


def patchify(fr: torch.Tensor, geo: Geometry, dtype=torch.float32):
    """uint8 [T, 3, H, W] -> (pixel_values [gt*gh*gw, C*Tp*P*P], grid (gt, gh, gw)).

    Same arithmetic and ordering as the HF Qwen2-VL video processor: fused rescale 1/255 +
    CLIP mean/std, then (gt, Tp, C, gh/m, m, P, gw/m, m, P) -> (gt, gh/m, gw/m, m, m, C, Tp, P, P).
    """
    t, c, h, w = fr.shape
    p, tp, m = geo.patch_size, geo.temporal_patch_size, geo.spatial_merge_size
    gt, gh, gw = t // tp, h // p, w // p
    # fused rescale + normalise exactly as HF's fast processor does it (bit-exact against Qwen2VLVideoProcessor 5.5.0,
    # tests/golden/frame_ingest.npz): mean' = mean * 255, std' = std * 255 in fp32, then (x - mean') / std'
    inv_rescale = 1.0 / (1.0 / 255.0)
    mean = (torch.tensor(CLIP_MEAN) * inv_rescale).to(fr.device).view(1, 3, 1, 1)
    std = (torch.tensor(CLIP_STD) * inv_rescale).to(fr.device).view(1, 3, 1, 1)
    x = (fr.to(torch.float32) - mean) / std
    x = x.reshape(gt, tp, c, gh // m, m, p, gw // m, m, p).permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return x.reshape(gt * gh * gw, c * tp * p * p).to(dtype).contiguous(), (gt, gh, gw)


def weight_shapes(geo: Geometry):
    """(HF parameter name, shape, kind) for every tensor on the path."""
    out = []
    v, e = "model.visual.", geo.vit_hidden
    out.append((v + "patch_embed.proj.weight", (e, geo.in_channels, geo.temporal_patch_size, geo.patch_size, geo.patch_size), "w"))
    for li in range(geo.vit_depth):
        b = f"{v}blocks.{li}."
        out += [(b + "norm1.weight", (e,), "n"), (b + "norm2.weight", (e,), "n"),
                (b + "attn.qkv.weight", (3 * e, e), "w"), (b + "attn.qkv.bias", (3 * e,), "b"),
                (b + "attn.proj.weight", (e, e), "w"), (b + "attn.proj.bias", (e,), "b"),
                (b + "mlp.gate_proj.weight", (geo.vit_mlp, e), "w"), (b + "mlp.gate_proj.bias", (geo.vit_mlp,), "b"),
                (b + "mlp.up_proj.weight", (geo.vit_mlp, e), "w"), (b + "mlp.up_proj.bias", (geo.vit_mlp,), "b"),
                (b + "mlp.down_proj.weight", (e, geo.vit_mlp), "w"), (b + "mlp.down_proj.bias", (e,), "b")]
    u = geo.spatial_merge_size ** 2
    out += [(v + "merger.ln_q.weight", (e,), "n"),
            (v + "merger.mlp.0.weight", (e * u, e * u), "w"), (v + "merger.mlp.0.bias", (e * u,), "b"),
            (v + "merger.mlp.2.weight", (geo.hidden, e * u), "w"), (v + "merger.mlp.2.bias", (geo.hidden,), "b")]
    lm, hd = "model.language_model.", geo.head_dim
    out.append((lm + "embed_tokens.weight", (geo.vocab, geo.hidden), "w"))
    for li in range(geo.num_layers):
        b = f"{lm}layers.{li}."
        out += [(b + "input_layernorm.weight", (geo.hidden,), "n"), (b + "post_attention_layernorm.weight", (geo.hidden,), "n"),
                (b + "self_attn.q_proj.weight", (geo.num_heads * hd, geo.hidden), "w"), (b + "self_attn.q_proj.bias", (geo.num_heads * hd,), "b"),
                (b + "self_attn.k_proj.weight", (geo.num_kv_heads * hd, geo.hidden), "w"), (b + "self_attn.k_proj.bias", (geo.num_kv_heads * hd,), "b"),
                (b + "self_attn.v_proj.weight", (geo.num_kv_heads * hd, geo.hidden), "w"), (b + "self_attn.v_proj.bias", (geo.num_kv_heads * hd,), "b"),
                (b + "self_attn.o_proj.weight", (geo.hidden, geo.num_heads * hd), "w"),
                (b + "mlp.gate_proj.weight", (geo.mlp, geo.hidden), "w"), (b + "mlp.up_proj.weight", (geo.mlp, geo.hidden), "w"),
                (b + "mlp.down_proj.weight", (geo.hidden, geo.mlp), "w")]
    out += [(lm + "norm.weight", (geo.hidden,), "n"), ("lm_head.weight", (geo.vocab, geo.hidden), "w")]
    return out


def random_weights(geo: Geometry, seed: int = 0, device="cpu", dtype=torch.bfloat16, std: float = 0.02):
    """Random-init state dict.  On CPU the values depend only on (geo, seed); on CUDA the
    device generator is used (fast, for the 7B / 32B benchmark geometries)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    w = {}
    for name, shape, kind in weight_shapes(geo):
        if kind == "n":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
        elif kind == "b":
            t = std * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
        else:
            fan_in = math.prod(shape[1:])
            s = std if "embed_tokens" in name else min(std * 2.0, 1.0 / math.sqrt(fan_in))
            t = torch.empty(shape, device=dev, dtype=dtype).normal_(0.0, s, generator=g) if dev.type == "cuda" \
                else (s * torch.randn(shape, generator=g, dtype=torch.float32))
        w[name] = t.to(dtype)
    return w


def block_weights_cpu(geo: Geometry, seed: int = 0, dtype=torch.bfloat16, layers=None, vit_blocks=None, block_elems=1 << 26):
    """CPU state dict for the big geometries in seconds instead of minutes: one seeded normal(0,1) block is
    generated once and every tensor is a scaled COPY of a different window of it (distinct memory, realistic
    values).  `layers` / `vit_blocks` restrict the dict to the first n LLM layers / ViT blocks."""
    g = torch.Generator().manual_seed(seed)
    block = torch.randn(block_elems, generator=g, dtype=torch.float32).to(dtype)
    w, off = {}, 0
    for name, shape, kind in weight_shapes(geo):
        if layers is not None and ".language_model.layers." in name and int(name.split(".layers.")[1].split(".")[0]) >= layers:
            continue
        if vit_blocks is not None and ".visual.blocks." in name and int(name.split(".blocks.")[1].split(".")[0]) >= vit_blocks:
            continue
        n = math.prod(shape)
        t = torch.empty(n, dtype=dtype)
        done = 0
        while done < n:
            take = min(n - done, block_elems - off)
            t[done:done + take] = block[off:off + take]
            done += take
            off = (off + take) % block_elems
        if kind == "n":
            t = 1.0 + 0.1 * t
        else:
            s = 0.02 if (kind == "b" or "embed_tokens" in name) else min(0.04, 1.0 / math.sqrt(math.prod(shape[1:])))
            t = t * s
        w[name] = t.to(dtype).reshape(shape)
    return w


def fast_forward(rounds: int, window_size=16, text_round=16, text_sink=512, text_sliding_window=512, n_vision_tokens=256,
                 previous_text_tokens=24, new_tokens=20, grid=(1, 32, 32), seed=0):
    """Host-only replay of `rounds` chunks of session bookkeeping (ids, history, eviction planning) with
    synthetic generated text, to put a session into its steady state without running the model.  Returns the
    state a session object needs: prev_ids (incl. the forced <|im_end|>), history, clips, pixels, grids, i, and
    cache_len (rows the cache holds for prev_ids)."""
    from .inference.inference import plan_past_kv
    hist, clips, pixels, grids = [], [], [], []
    prev, cache_len = None, 0
    for i in range(rounds):
        if prev is not None:
            plan = plan_past_kv(prev, i, text_round, window_size, hist, 3, 2, clips, pixels, text_sink, text_sliding_window)
            cache_len = sum(r[2] for r in plan.runs(cache_len))
            prev = plan.ids.tolist()
        clips.append(None)
        if i == 0:
            hist[:] = [{"role": "previous text", "content": ""},
                       {"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}, {"type": "text", "text": "query"}]}]
        else:
            hist.append({"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}]})
        prompt = round_prompt(i, n_vision_tokens, previous_text_tokens, 8, seed)
        ids = prompt if prev is None else prev + (prompt if prev[-1] != LF else prompt[1:])
        pixels.append(None)
        grids.append(tuple(grid))
        gen = _text_tokens(new_tokens, seed * 7919 + 31 * i + 5)
        seq = ids + gen
        cache_len = len(seq) - 1
        seq = seq + [IM_END]
        prev = seq
        hist.append({"role": "assistant", "content": " ".join(str(t) for t in gen + [IM_END]) + " ..."})
    return {"prev_ids": prev, "history": hist, "clips": clips, "pixels": pixels, "grids": grids, "i": rounds,
            "cache_len": cache_len}
