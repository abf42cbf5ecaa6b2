"""`model.generate` for the streaming session (greedy / temperature sampling only).

The reference copies HF's `generate` / `_sample` (`generate/streaming_generate_qwen.py:8-127, 129-278`) to
inject its cache class and to assert greedy-or-sample mode (:252).  This is the same loop reduced to what
the session uses: one prefill forward over the uncached ids (with pixel values), then one forward per new
token without them (`generate/prepare_generation.py:131-134`), repetition penalty and temperature as logits
processors, stop on eos or `max_new_tokens`, returning `.sequences` and `.past_key_values`.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from ..generate.streaming_cache import StreamingCache


@torch.no_grad()
def streaming_generate(self, input_ids=None, attention_mask=None, pixel_values_videos=None, video_grid_thw=None,
                       second_per_grid_ts=None, past_key_values=None, max_new_tokens=20, use_cache=True,
                       return_dict_in_generate=True, do_sample=False, repetition_penalty=1.0, temperature=1.0, streaming_args=None,
                       pad_token_id=151645, eos_token_id=151645, suppress_tokens=None, **kwargs):
    if streaming_args is None:
        raise RuntimeError("streaming_args is required")
    if past_key_values is None:
        from ...config import from_hf_config
        past_key_values = StreamingCache(from_hf_config(self.config), capacity=kwargs.pop("cache_capacity", 8192), device=input_ids.device)
    cache = past_key_values
    seq = input_ids
    new = 0
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
        seq = torch.cat([seq, nxt], dim=-1)
        new += 1
        if int(nxt) == eos_token_id or new >= max_new_tokens:
            break
    if return_dict_in_generate:
        return SimpleNamespace(sequences=seq, past_key_values=cache)
    return seq
