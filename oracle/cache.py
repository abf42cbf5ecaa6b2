"""Oracle: the reference's list-of-tensors KV cache (test infrastructure, CPU only).

Restates `StreamingCache` (`generate/streaming_cache.py:6-74`) on top of the HF 4.52
`DynamicCache` surface the reference touches: `key_cache[i]`, `value_cache[i]`
(`[1, H_kv, S, D]`, grown with `torch.cat` on dim -2, `:72-73`), `_seen_tokens`
(`:54-55`), iteration yielding `(K, V)` per layer, `get_seq_length()`, and the stored
position ids (`update_position_ids`, `:10-28`).  In shrink mode the cache holds the RAW
(un-rotated) keys (`qwen2_5/language_forward.py:190-192`).
"""
from __future__ import annotations

import torch


class ListCache:
    def __init__(self):
        self.key_cache = []
        self.value_cache = []
        self.position_ids_cache = []
        self._seen_tokens = 0

    def __iter__(self):
        for k, v in zip(self.key_cache, self.value_cache):
            yield k, v

    def __len__(self):
        return len(self.key_cache)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        if len(self.key_cache) <= layer_idx or self.key_cache[layer_idx].numel() == 0:
            return 0
        return self.key_cache[layer_idx].shape[-2]

    def update_position_ids(self, position_ids, layer_idx: int = 0):
        if position_ids is not None:
            while len(self.position_ids_cache) < layer_idx:
                self.position_ids_cache.append(torch.tensor([]))
            if len(self.position_ids_cache) == layer_idx:
                self.position_ids_cache.append(position_ids)
            else:
                self.position_ids_cache[layer_idx] = position_ids
        return self.position_ids_cache[layer_idx]

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        if layer_idx == 0:
            self._seen_tokens += key_states.shape[-2]
        while len(self.key_cache) < layer_idx:
            self.key_cache.append(torch.tensor([]))
            self.value_cache.append(torch.tensor([]))
        if len(self.key_cache) == layer_idx:
            self.key_cache.append(key_states)
            self.value_cache.append(value_states)
        elif self.key_cache[layer_idx].numel() == 0:
            self.key_cache[layer_idx] = key_states
            self.value_cache[layer_idx] = value_states
        else:
            self.key_cache[layer_idx] = torch.cat([self.key_cache[layer_idx], key_states], dim=-2)
            self.value_cache[layer_idx] = torch.cat([self.value_cache[layer_idx], value_states], dim=-2)
        return self.key_cache[layer_idx], self.value_cache[layer_idx]
