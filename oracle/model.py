"""Oracle: model forward and the streaming session loop (test infrastructure, CPU only).

Restates, op for op and in the reference's order,
  * `model_forward` (`qwen2_5/model_forward.py:30-176`): token embedding, ViT on the new
    chunk, `masked_scatter` of the video embeddings (`:87-93`), shrink-mode positions from
    the FULL pruned ids (`:101-110`),
  * `streaming_language_model_forward` / `streaming_text_decoder_layer_forward` /
    `streaming_text_flash_attn_forward` (`qwen2_5/language_forward.py:143-305, 395-516`):
    raw K/V appended to the cache first (`:190-192`), cos/sin for ALL S positions
    (`:454`), every cached key re-rotated each forward (`:197`), pad of the id buffer
    by one 0 after each forward (`:505-507`),
  * lm_head (`model_forward.py:316-317`; only the last position is evaluated here -- the
    reference computes all of them but samples from `[:, -1]`,
    `generate/streaming_generate_qwen.py:73`),
  * the greedy branch of `_sample` (`streaming_generate_qwen.py:53-110`),
  * the per-chunk session loop of `streaming_inference` (`inference/inference.py:309-517`):
    eviction first (`:319`), id concatenation with the "\\n" rule (`:397-406`), growing
    `video_grid_thw` (`:409-416`), forced `<|im_end|>` (`:456-459`).

Weights: dict of tensors with HF parameter names.  Geometry: plain dict (see
`streaming-vlm_b200/config.py::Geometry.as_dict`).
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F

from .attention import attention
from .cache import ListCache
from .eviction import process_past_kv
from .positions import append_mode_positions, apply_mrope, default_inv_freq, get_rope_index, rotary_cos_sin
from .tokens import IM_END, LF, VIDEO_PAD
from .vit import rmsnorm, vision_forward


def llm_layer(x, w, li, geo, cache, cos, sin, taps=None, pos_mode="shrink"):
    p = f"model.language_model.layers.{li}."
    t = x.shape[0]
    hq, hkv, d = geo["num_heads"], geo["num_kv_heads"], geo["head_dim"]
    h = rmsnorm(x, w[p + "input_layernorm.weight"], geo["rms_eps"])
    q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"])
    k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"])
    v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
    q = q.view(1, t, hq, d).transpose(1, 2)
    k = k.view(1, t, hkv, d).transpose(1, 2)
    v = v.view(1, t, hkv, d).transpose(1, 2)
    if pos_mode == "append":
        # append mode (language_forward.py:184-188): cos/sin cover the NEW tokens only; the ROTATED keys are cached and
        # never touched again
        q_rot, k_new = apply_mrope(q, k, cos, sin, geo["mrope_section"])
        k_rot, v_all = cache.update(k_new, v, li)
    else:
        k_all, v_all = cache.update(k, v, li)                  # raw keys are cached
        q_rot, k_rot = apply_mrope(q, k_all, cos, sin, geo["mrope_section"])
    o = attention(q_rot[0], k_rot[0], v_all[0], causal=True)
    if taps is not None:
        taps.setdefault("attn_out", []).append(o)
        taps.setdefault("k_rot", []).append(k_rot[0])
        taps.setdefault("q_rot", []).append(q_rot[0])
    x = x + F.linear(o, w[p + "self_attn.o_proj.weight"])
    h = rmsnorm(x, w[p + "post_attention_layernorm.weight"], geo["rms_eps"])
    g = F.linear(h, w[p + "mlp.gate_proj.weight"])
    u = F.linear(h, w[p + "mlp.up_proj.weight"])
    return x + F.linear(F.silu(g) * u, w[p + "mlp.down_proj.weight"])


def model_forward(new_ids, full_ids, grids, cache, w, geo, pixels=None, pixel_grid=None, taps=None, timers=None,
                  pos_mode="shrink", state=None):
    """One forward over the uncached ids.  Returns fp32 logits of the last position.
    `pos_mode="append"` (model_forward.py:112-153): positions of the new tokens only, continuing after
    `state["last_cache_position"]`; the grids of THIS call (`pixel_grid`) feed `get_rope_index`."""
    ids_t = torch.tensor(new_ids, dtype=torch.long)
    x = F.embedding(ids_t, w["model.language_model.embed_tokens.weight"])
    if pixels is not None:
        t0 = time.perf_counter()
        vid = vision_forward(pixels, pixel_grid, w, geo, timers=timers)
        if timers is not None:
            timers["vit_total"] = timers.get("vit_total", 0.0) + time.perf_counter() - t0
        mask = ids_t == VIDEO_PAD
        if int(mask.sum()) != vid.shape[0]:
            raise ValueError(f"Video features and video tokens do not match: tokens: {int(mask.sum())}, features {vid.shape[0]}")
        x = x.clone()
        x[mask] = vid.to(x.dtype)
        if taps is not None:
            taps["video_embeds"] = vid
    if pos_mode == "append":
        pos = append_mode_positions(new_ids, pixel_grid, cache.get_seq_length(), state, geo["spatial_merge_size"], geo["tokens_per_second"])
    else:
        pos, _ = get_rope_index(full_ids, grids, geo["spatial_merge_size"], geo["tokens_per_second"])
        cache.update_position_ids(pos)
    cos, sin = rotary_cos_sin(pos, default_inv_freq(geo["head_dim"], geo["rope_theta"]), x.dtype)
    if taps is not None:
        taps["position_ids"] = pos
    t0 = time.perf_counter()
    for li in range(geo["num_layers"]):
        x = llm_layer(x, w, li, geo, cache, cos, sin, taps, pos_mode)
    if timers is not None:
        key = "prefill_layers" if len(new_ids) > 1 else "decode_layers"
        timers[key] = timers.get(key, 0.0) + time.perf_counter() - t0
        timers["n_" + key] = timers.get("n_" + key, 0) + 1
    x = rmsnorm(x, w["model.language_model.norm.weight"], geo["rms_eps"])
    logits = F.linear(x[-1:], w["lm_head.weight"]).float()[0]
    if taps is not None:
        taps["logits"] = logits
    return logits


class Session:
    """The reference's per-stream state + chunk loop, with synthetic prompts supplied by
    the caller as token-id lists (no tokenizer exists offline)."""

    def __init__(self, w, geo, window_size=16, text_round=16, text_sink=512, text_sliding_window=512,
                 max_new_tokens=20, suppress=None, assistant_start_bias=3, assistant_end_bias=2, pos_mode="shrink",
                 repetition_penalty=1.0, temperature=1.0):
        assert pos_mode in ("append", "shrink"), "pos_mode must be in ['append', 'shrink']"      # streaming_args.py:5
        self.pos_mode = pos_mode
        # logits processors of the reference's generate call (inference.py:41-42, 433-437): HF RepetitionPenaltyLogitsProcessor
        # over the whole `input_ids` so far, then TemperatureLogitsWarper (streaming_generate_qwen.py:75); token selection here
        # stays the argmax of the processed scores (the multinomial draw is not reproducible across RNG implementations)
        self.repetition_penalty, self.temperature = float(repetition_penalty), float(temperature)
        self.pos_state = {"last_cache_position": -1}
        self.w, self.geo = w, geo
        self.window_size, self.text_round = window_size, text_round
        self.text_sink, self.text_sliding_window = text_sink, text_sliding_window
        self.max_new_tokens = max_new_tokens
        self.suppress = suppress          # bool [vocab] of ids never emitted (synthetic-stream guard)
        self.asb, self.aeb = assistant_start_bias, assistant_end_bias
        self.cache = None
        self.prev_ids = None              # tensor [1, S]
        self.history = []
        self.clips, self.pixels = [], []
        self.grids = []                   # grows forever, as in the reference (:415)
        self.i = 0
        self.timing = []

    def step(self, prompt_ids, pixels, grid, previous_text="", taps=None):
        """One chunk.  prompt_ids: the tokenised new round (starting with "\\n" for i>0)."""
        i = self.i
        t0 = time.perf_counter()
        res = process_past_kv(self.cache, i, self.text_round, self.window_size, self.history, self.prev_ids,
                              self.asb, self.aeb, self.clips, self.pixels, self.text_sink,
                              self.text_sliding_window, return_gather=True)
        self.cache, self.prev_ids, self.clips, self.pixels, gather = res
        t_pkv = time.perf_counter() - t0
        self.clips.append(None)
        if i == 0:
            self.history = [{"role": "previous text", "content": previous_text},
                            {"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"},
                                                         {"type": "text", "text": "query"}]}]
        else:
            self.history.append({"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}]})
        prompt = [int(t) for t in prompt_ids]
        if self.prev_ids is not None:
            prev = self.prev_ids.flatten().tolist()
            ids = prev + (prompt if prev[-1] != LF else prompt[1:])
        else:
            ids = prompt
        self.pixels.append(pixels)
        self.grids.append(tuple(grid))
        if self.cache is None:
            self.cache = ListCache()
        input_len = len(ids)

        t0 = time.perf_counter()
        timers = {}
        full = list(ids)                       # streaming_args.input_ids
        seq = list(ids)
        n_fwd = 0
        step_logits = []
        while True:
            cached = self.cache.get_seq_length()
            new = seq[cached:]
            px = pixels if n_fwd == 0 else None
            logits = model_forward(new, full, self.grids, self.cache, self.w, self.geo, px, [tuple(grid)],
                                   taps if n_fwd == 0 else None, timers, self.pos_mode, self.pos_state)
            full = full + [0]                   # F.pad(ids, (0,1)) after every forward
            n_fwd += 1
            if self.repetition_penalty != 1.0:
                seen = torch.tensor(seq, dtype=torch.long)
                score = logits[seen]
                logits = logits.scatter(0, seen, torch.where(score < 0, score * self.repetition_penalty, score / self.repetition_penalty))
            if self.temperature != 1.0:
                logits = logits / self.temperature
            if self.suppress is not None:
                logits = logits.masked_fill(self.suppress, float("-inf"))
            step_logits.append(logits)
            tok = int(torch.argmax(logits))
            seq.append(tok)
            if tok == IM_END or len(seq) - input_len >= self.max_new_tokens:
                break
        t_gen = time.perf_counter() - t0
        if seq[-1] != IM_END:
            seq.append(IM_END)
        new_tokens = seq[input_len:]
        self.prev_ids = torch.tensor([seq], dtype=torch.long)
        self.history.append({"role": "assistant", "content": " ".join(str(t) for t in new_tokens) + " ..."})
        self.timing.append({"PKV": t_pkv, "GEN": t_gen, "n_forward": n_fwd, "new_tokens": len(new_tokens), **timers})
        self.i += 1
        return {"new_tokens": new_tokens, "input_len": input_len, "ids": seq, "gather": gather,
                "cache_len": self.cache.get_seq_length(), "step_logits": step_logits}
