"""`model.generate` for the streaming session.

The reference copies HF's `generate` / `_sample` (`generate/streaming_generate_qwen.py:8-127, 129-278`) to inject its
cache class and to assert greedy-or-sample mode (:252).  Here `generate` keeps that call signature -- the session loop
calls `model.generate(**inputs, past_key_values=..., max_new_tokens=20, use_cache=True, return_dict_in_generate=True,
do_sample=True, repetition_penalty=1.05, streaming_args=..., pad_token_id=151645, temperature=0.9)`
(`inference/inference.py:440-451`) and reads `.sequences` / `.past_key_values` -- and has two executions:

* FAST PATH (Qwen2.5-VL, the default): the whole turn runs on the session runner built over the live model's own
  parameters (`StreamingEngine.from_hf_model`, no second copy of the weights): ViT as one CUDA graph, chunk prefill as
  one CUDA graph, one CUDA-graph replay per decoded token with look-ahead, token selection (repetition penalty,
  temperature, sampling) inside the lm_head kernel.  The reference's unchanged loop therefore gets the engine's speed.
* EAGER PATH (`SVLM_DROPIN_EAGER=1`, or a Qwen2-VL model): one call of the patched HF forwards per step -- every
  MethodType patch point of `qwen2_5/patch_model.py` executes -- with the logits processors in torch.

Both modes of `StreamingArgs.pos_mode` ("shrink": contiguous positions of the full pruned ids every turn; "append":
positions continue after `last_cache_position`) and `all_text` are supported on both paths.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from ..generate.streaming_cache import StreamingCache


def _fast_path_engine(model, cache, capacity):
    """The session runner bound to this model (built on first use) with `cache` adopted as its arena."""
    from ...engine import StreamingEngine
    eng = getattr(model, "_svlm_engine", None)
    if eng is None:
        eng = StreamingEngine.from_hf_model(model, capacity=capacity, cache=cache)
        object.__setattr__(model, "_svlm_engine", eng)
    if cache is not None and eng.cache is not cache:
        eng.adopt_cache(cache)
    return eng


def _use_fast_path(model) -> bool:
    if os.environ.get("SVLM_DROPIN_EAGER", "0") == "1":
        return False
    return getattr(model.config, "model_type", "") == "qwen2_5_vl" and getattr(model, "_svlm_fast_path", True)


@torch.no_grad()
def streaming_generate(self, input_ids=None, attention_mask=None, pixel_values_videos=None, video_grid_thw=None,
                       second_per_grid_ts=None, past_key_values=None, max_new_tokens=20, use_cache=True,
                       return_dict_in_generate=True, do_sample=False, repetition_penalty=1.0, temperature=1.0, streaming_args=None,
                       pad_token_id=151645, eos_token_id=151645, suppress_tokens=None, seed=None, forced_tokens=None, **kwargs):
    """`forced_tokens` (test hook, the purpose of the reference's --gt_json mode, inference.py:231-239): token j of the turn is
    replaced by `forced_tokens[j]` after it was selected; the model's own choices stay readable in `model._svlm_last_selected`."""
    if streaming_args is None:
        raise RuntimeError("streaming_args is required")
    if input_ids is None or input_ids.dim() != 2 or input_ids.shape[0] != 1:
        raise RuntimeError("batch size 1 only (the streaming path is single-stream)")
    capacity = kwargs.pop("cache_capacity", 8192)
    if not _use_fast_path(self):
        return _eager_generate(self, input_ids, pixel_values_videos, video_grid_thw, past_key_values, max_new_tokens, return_dict_in_generate,
                               do_sample, repetition_penalty, temperature, streaming_args, eos_token_id, suppress_tokens, capacity,
                               forced_tokens)
    from ...engine import GenConfig
    eng = _fast_path_engine(self, past_key_values, capacity)
    cache = eng.cache
    if suppress_tokens is not None:
        eng.set_suppress(suppress_tokens)
    if cache.get_seq_length() == 0:
        cache.pos_mode = streaming_args.pos_mode
    elif cache.pos_mode != streaming_args.pos_mode:
        raise RuntimeError(f"the cache was filled in pos_mode={cache.pos_mode!r}; streaming_args says {streaming_args.pos_mode!r}")
    n_in = input_ids.shape[1]
    grid = None
    if pixel_values_videos is not None:
        if video_grid_thw is None or video_grid_thw.shape[0] != 1:
            raise RuntimeError("one video chunk per generate call (video_grid_thw [1, 3])")
        grid = tuple(int(v) for v in video_grid_thw[0].tolist())
    if streaming_args.pos_mode == "shrink":
        # contiguous positions of the FULL pruned sequence (reference model_forward.py:101-110)
        full = streaming_args.input_ids if streaming_args.input_ids is not None else input_ids
        if streaming_args.video_grid_thw is not None:
            cache.set_video_grids(streaming_args.video_grid_thw)
        cache.compute_positions(full, all_text=streaming_args.all_text)
    else:
        # positions of the uncached suffix only, continuing after last_cache_position (reference model_forward.py:112-153)
        cache.reserve_positions(n_in - cache.get_seq_length() + max_new_tokens)
        cache.append_positions(input_ids, cache.get_seq_length(), video_grid_thw if pixel_values_videos is not None else None,
                               all_text=streaming_args.all_text)
    ids_host = input_ids[0].cpu().numpy()
    gen = GenConfig(do_sample=bool(do_sample), repetition_penalty=float(repetition_penalty), temperature=float(temperature) if do_sample else 1.0,
                    seed=None if seed is None else int(seed))
    seq, n_fwd = eng.generate(ids_host, pixel_values_videos, grid, forced=forced_tokens, gen=gen, max_new_tokens=max_new_tokens, eos=eos_token_id)
    object.__setattr__(self, "_svlm_last_selected", list(eng.last_argmax))
    if streaming_args.pos_mode == "append":
        cache.record_decode_positions(n_in)
        streaming_args.last_cache_position = int(cache.next_pos.item()) - 1
    elif streaming_args.input_ids is not None:
        # one pad per forward keeps `streaming_args.input_ids` as long as the cache (reference language_forward.py:505-507)
        streaming_args.input_ids = F.pad(streaming_args.input_ids, (0, n_fwd), "constant", 0)
    new = torch.tensor([seq[n_in:]], dtype=input_ids.dtype, device=input_ids.device)
    sequences = torch.cat([input_ids, new], dim=1)
    if return_dict_in_generate:
        return SimpleNamespace(sequences=sequences, past_key_values=cache)
    return sequences


def _eager_generate(self, input_ids, pixel_values_videos, video_grid_thw, past_key_values, max_new_tokens, return_dict_in_generate,
                    do_sample, repetition_penalty, temperature, streaming_args, eos_token_id, suppress_tokens, capacity, forced_tokens):
    """One patched-forward call per step (the reference's `_sample` loop, `generate/streaming_generate_qwen.py:53-110`): a
    prefill forward over the uncached ids with the pixel values, then one forward per new token without them
    (`generate/prepare_generation.py:131-134`); repetition penalty and temperature as logits processors (:73-76), multinomial or
    argmax (:96-99), stop on eos or `max_new_tokens`."""
    if past_key_values is None:
        from ...config import from_hf_config
        past_key_values = StreamingCache(from_hf_config(self.config), capacity=capacity, device=input_ids.device, pos_mode=streaming_args.pos_mode)
    cache = past_key_values
    if cache.get_seq_length() == 0:
        cache.pos_mode = streaming_args.pos_mode
    seq = input_ids
    new = 0
    selected = []
    while True:
        cached = cache.get_seq_length()
        step_ids = seq[:, cached:]
        first = new == 0
        out = self(input_ids=step_ids, pixel_values_videos=pixel_values_videos if first else None,
                   video_grid_thw=video_grid_thw if first else None, past_key_values=cache, streaming_args=streaming_args,
                   logits_to_keep=1)
        logits = out.logits[:, -1, :].float()
        if repetition_penalty != 1.0:
            score = torch.gather(logits, 1, seq)
            score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
            logits = logits.scatter(1, seq, score)
        if suppress_tokens is not None:
            logits = logits.masked_fill(suppress_tokens, float("-inf"))
        if do_sample:
            probs = torch.softmax(logits / temperature, dim=-1)
            nxt = torch.multinomial(probs, num_samples=1)
        else:
            nxt = torch.argmax(logits, dim=-1, keepdim=True)
        selected.append(int(nxt))
        if forced_tokens is not None and new < len(forced_tokens):
            nxt = torch.full_like(nxt, int(forced_tokens[new]))
        seq = torch.cat([seq, nxt], dim=-1)
        new += 1
        if int(nxt) == eos_token_id or new >= max_new_tokens:
            break
    object.__setattr__(self, "_svlm_last_selected", selected)
    if streaming_args.pos_mode == "append":
        streaming_args.last_cache_position = int(cache.next_pos.item()) - 1
    if return_dict_in_generate:
        return SimpleNamespace(sequences=seq, past_key_values=cache)
    return seq
