"""StreamingCache backed by a fixed-capacity HBM arena.

Interface of the reference's `StreamingCache` (`generate/streaming_cache.py:6-74`) -- `update`,
`update_position_ids`, `get_seq_length`, iteration over `(K, V)` per layer, list-like
`key_cache[i]` / `value_cache[i]` that the session code reads and assigns -- on top of ONE
pre-allocated bf16 tensor `[2 buffers][3 planes][L][H_kv][cap][D]` (K_raw, K_rot, V; see
`include/svlm_b200.h`).  The reference re-allocates and copies every layer's cache on every forward
(`torch.cat`, `:72-73`); here a forward appends rows in place (`svlm_qkv_rope_append`) and an
eviction is a single ping-pong compaction launch (`svlm_kv_compact_rope`).

Sized for a B200: the 7B geometry at `capacity=8192` takes 2*3*28*4*8192*128*2 B = 1.4 GB; a
65 536-token arena 11.3 GB of the 180 GB.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _abi
from ...config import Geometry
from ...utils.get_qwen_range import TOKEN_IDS, as_host_ids

K_RAW, K_ROT, V_PLANE = 0, 1, 2


def rope_tables(geo: Geometry, rows: int, device):
    """cos/sin [rows, D/2] bf16 for integer positions: fp32 `p * inv_freq`, fp32 cos/sin, one rounding --
    the arithmetic of HF `Qwen2_5_VLRotaryEmbedding.forward` (called at qwen2_5/language_forward.py:454)."""
    inv_freq = 1.0 / (geo.rope_theta ** (torch.arange(0, geo.head_dim, 2, dtype=torch.int64).to(torch.float32) / geo.head_dim))
    freqs = torch.arange(rows, dtype=torch.float32)[:, None] * inv_freq[None, :]
    return freqs.cos().to(torch.bfloat16).to(device).contiguous(), freqs.sin().to(torch.bfloat16).to(device).contiguous()


class _PlaneList:
    """`cache.key_cache[i]` / `cache.value_cache[i]`: views `[1, H, S, D]` of the live buffer; assignment
    copies into the arena (what `prune_id_and_kv_cache` of the reference does, inference.py:58-59)."""

    def __init__(self, cache, plane):
        self._c, self._plane = cache, plane

    def __len__(self):
        return self._c.layers if self._c.seq_len > 0 else 0

    def __getitem__(self, layer):
        c = self._c
        return c.arena[c.cur, self._plane, layer, :, :c.seq_len].unsqueeze(0)

    def __setitem__(self, layer, value):
        c = self._c
        n = value.shape[-2]
        if n > c.capacity:
            raise RuntimeError(f"StreamingCache capacity {c.capacity} exceeded ({n})")
        c.arena[c.cur, self._plane, layer, :, :n].copy_(value.reshape(c.kv_heads, n, c.head_dim))
        if self._plane == K_RAW:
            c._rot_stale = True
        if layer == c.layers - 1 and self._plane == V_PLANE:
            c.seq_len = n


class StreamingCache:
    def __init__(self, geo: Geometry = None, capacity: int = 8192, device="cuda", max_position: int = None,
                 _distributed_cache_data=None):
        if geo is None:
            raise RuntimeError("StreamingCache(geo=...) needs the model geometry to size its arena")
        if torch.device(device).type != "cuda":
            raise RuntimeError("StreamingCache lives in HBM; there is no CPU fallback")
        _abi.lib()
        self.geo = geo
        self.layers, self.kv_heads, self.head_dim = geo.num_layers, geo.num_kv_heads, geo.head_dim
        self.capacity = int(capacity)
        self.device = torch.device(device)
        # The TMA / UMMA tiles want a 1024-byte aligned base; torch's caching allocator only promises 512 bytes.
        shape = (2, 3, self.layers, self.kv_heads, self.capacity, self.head_dim)
        n_el = 1
        for d in shape:
            n_el *= d
        self._arena_storage = torch.zeros(n_el + 512, dtype=torch.bfloat16, device=self.device)
        skip = (-self._arena_storage.data_ptr() % 1024) // 2
        self.arena = self._arena_storage[skip:skip + n_el].view(shape)
        self.view = _abi.ArenaView(self.arena)
        self.cur = 0
        self.seq_len = 0
        self._seen_tokens = 0
        self.section = tuple(geo.mrope_section)
        self.cos_table, self.sin_table = rope_tables(geo, max_position or self.capacity, self.device)
        self.pos = torch.zeros(3, self.capacity, dtype=torch.int32, device=self.device)
        self.next_pos = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ids_dev = torch.zeros(self.capacity, dtype=torch.int64, device=self.device)
        self.ids_pinned = torch.zeros(self.capacity, dtype=torch.int64).pin_memory()
        self.grids = torch.tensor([[1, 32, 32]], dtype=torch.int32, device=self.device)
        self.n_grids = 1
        self.spatial_merge = geo.spatial_merge_size
        self.t_step = 2 * geo.tokens_per_second // 2          # (2 / FPS) * tokens_per_second with FPS = 2
        self.position_ids_cache = []
        self.key_cache = _PlaneList(self, K_RAW)
        self.value_cache = _PlaneList(self, V_PLANE)
        self._rot_stale = False
        self.compactions = 0
        self.decode_splits = max(1, min((self.capacity + 127) // 128, 296 // self.kv_heads))
        self._pending_start = 0

    # ------------------------------------------------------------------ reference interface
    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.seq_len

    def __len__(self):
        return self.layers if self.seq_len > 0 else 0

    def __iter__(self):
        for li in range(self.layers if self.seq_len > 0 else 0):
            yield self.key_cache[li], self.value_cache[li]

    def update_position_ids(self, position_ids, layer_idx: int = 0):
        if position_ids is not None:
            while len(self.position_ids_cache) < layer_idx:
                self.position_ids_cache.append(torch.tensor([]))
            if len(self.position_ids_cache) == layer_idx:
                self.position_ids_cache.append(position_ids)
            else:
                self.position_ids_cache[layer_idx] = position_ids
        return self.position_ids_cache[layer_idx]

    def layer_start(self, layer_idx: int, t: int) -> int:
        """Arena row at which layer `layer_idx` appends `t` new tokens.  All layers of one forward append at the
        same row; the sequence length advances once the last layer's append has been enqueued."""
        if layer_idx == 0:
            self._pending_start = self.seq_len
            self._seen_tokens += t
        start = self._pending_start
        if start + t > self.capacity:
            raise RuntimeError(f"StreamingCache capacity {self.capacity} exceeded ({start + t})")
        if layer_idx == self.layers - 1:
            self.seq_len = start + t
        return start

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        """Append raw K / V `[1, H, T, D]` for one layer; returns the `[1, H, S, D]` views (reference :30-74).
        K_rot for the new rows is NOT produced here (the attention op does that, `svlm_qkv_rope_append`);
        this method exists for callers that drive the cache the HF way."""
        t = key_states.shape[-2]
        start = self.layer_start(layer_idx, t)
        self.arena[self.cur, K_RAW, layer_idx, :, start:start + t].copy_(key_states.reshape(self.kv_heads, t, self.head_dim))
        self.arena[self.cur, V_PLANE, layer_idx, :, start:start + t].copy_(value_states.reshape(self.kv_heads, t, self.head_dim))
        self._rot_stale = True
        n = start + t
        return (self.arena[self.cur, K_RAW, layer_idx, :, :n].unsqueeze(0), self.arena[self.cur, V_PLANE, layer_idx, :, :n].unsqueeze(0))

    # ------------------------------------------------------------------ B200 path
    def set_video_grids(self, grids):
        """grids: iterable of (t, h, w) consumed from index 0 by `compute_positions`, as the reference's
        ever-growing `streaming_args.video_grid_thw` is (inference.py:415)."""
        g = torch.as_tensor(np.asarray(grids, dtype=np.int32).reshape(-1, 3))
        self.grids = g.to(self.device)
        self.n_grids = g.shape[0]

    def upload_ids(self, ids) -> int:
        if isinstance(ids, torch.Tensor) and ids.is_cuda:
            n = ids.numel()
            if n > self.capacity:
                raise RuntimeError(f"sequence of {n} tokens exceeds the arena capacity {self.capacity}")
            self.ids_dev[:n].copy_(ids.reshape(-1))
            return n
        ids = as_host_ids(ids)
        n = ids.shape[0]
        if n > self.capacity:
            raise RuntimeError(f"sequence of {n} tokens exceeds the arena capacity {self.capacity}")
        self.ids_pinned[:n] = torch.from_numpy(ids)
        self.ids_dev[:n].copy_(self.ids_pinned[:n], non_blocking=True)
        return n

    def compute_positions(self, ids) -> int:
        """Contiguous mRoPE positions of `ids` into `self.pos[:, :n]`; `self.next_pos` = max + 1."""
        n = self.upload_ids(ids)
        _abi.mrope_positions(self.ids_dev, n, self.grids, self.n_grids, self.pos, self.next_pos, self.status,
                             self.spatial_merge, self.t_step, TOKEN_IDS["<|video_pad|>"])
        return n

    # -- append position mode (reference: model_forward.py:112-153 + language_forward.py:184-188) -----------------
    def gather_positions(self, runs):
        """Retained tokens KEEP their positions in append mode: move them with the same runs as the K/V rows, so that the
        compaction kernel re-derives exactly the K_rot the token was appended with."""
        n = sum(r[2] for r in runs)
        self.pos[:, :n] = torch.cat([self.pos[:, s_:s_ + l] for s_, _, l in runs], dim=1)

    def append_positions(self, ids, start: int, grid) -> int:
        """Positions of the uncached ids `ids[start:]` only: `get_rope_index` over the new ids with THIS chunk's grid, shifted by
        last_cache_position + 1 (= `next_pos`, which the decode steps keep advancing); nothing is recomputed for the cache."""
        n = self.upload_ids(ids)
        n_new = n - start
        offset = self.next_pos.clone()
        if start == 0:
            offset.zero_()
        self.set_video_grids([grid])
        if not hasattr(self, "_pos_rel"):
            self._pos_rel = torch.zeros_like(self.pos)
        _abi.mrope_positions(self.ids_dev[start:], n_new, self.grids, self.n_grids, self._pos_rel, self.next_pos, self.status,
                             self.spatial_merge, self.t_step, TOKEN_IDS["<|video_pad|>"])
        self.pos[:, start:n] = self._pos_rel[:, :n_new] + offset
        self.next_pos += offset
        return n

    def compact(self, runs, new_ids=None):
        """Gather `runs` = [(src, dst, len)] from the live buffer into the other one, re-rotating K at the
        positions of `new_ids` (the pruned sequence; omitted = positions already in `self.pos`)."""
        if new_ids is not None:
            self.compute_positions(new_ids)
        n = sum(r[2] for r in runs)
        for lo in range(0, max(len(runs), 1), _abi.MAX_RUNS):   # > MAX_RUNS runs: several launches, same buffers
            part = runs[lo:lo + _abi.MAX_RUNS]
            if not part:
                break
            if lo > 0:
                raise RuntimeError(f"gather has {len(runs)} runs; SVLM_MAX_RUNS is {_abi.MAX_RUNS}")
            _abi.kv_compact_rope(self.view, self.cur, part, self.pos, self.cos_table, self.sin_table, self.section)
        self.cur ^= 1
        self.seq_len = n
        self._rot_stale = False
        self.compactions += 1

    def planes(self, layer):
        b = self.arena[self.cur]
        n = self.seq_len
        return b[K_RAW, layer, :, :n], b[K_ROT, layer, :, :n], b[V_PLANE, layer, :, :n]
