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
        """`cache.key_cache[i] = t` / `cache.value_cache[i] = t` with t `[1, H, n, D]`: rows 0..n-1 of that layer's plane are
        replaced.  K_rot of the layer is no longer the rotation of its K_raw, and the positions recorded for the old rows
        no longer describe the new ones: both are marked stale and re-derived before the next read (`ensure_rotated`)."""
        c = self._c
        n = value.shape[-2]
        if n > c.capacity:
            raise RuntimeError(f"StreamingCache capacity {c.capacity} exceeded ({n})")
        c.arena[c.cur, self._plane, layer, :, :n].copy_(value.reshape(c.kv_heads, n, c.head_dim))
        if self._plane == K_RAW:
            c._rot_stale = True
            c._pos_len = 0
        if layer == c.layers - 1 and self._plane == V_PLANE:
            c.seq_len = n


class StreamingCache:
    def __init__(self, geo: Geometry = None, capacity: int = 8192, device="cuda", max_position: int = None,
                 pos_mode: str = "shrink", _distributed_cache_data=None):
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
        # The kernels clamp a position to the last table row instead of reading beyond it, which would rotate with the wrong angle
        # silently: the host rules that out.  "shrink" positions never exceed the number of cached tokens (<= capacity), so the table
        # must cover the capacity; "append" positions grow for ever and are checked as they are handed out (reserve_positions).
        if max_position is not None and int(max_position) < self.capacity:
            raise ValueError(f"max_position {max_position} is smaller than the arena capacity {self.capacity}: contiguous positions "
                             "reach the capacity, the rotary tables must cover them")
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
        # K_rot bookkeeping: `_rot_stale` = some K_raw rows were written without their rotation (HF-style `update` before the
        # positions were known, or `key_cache[i] = ...`); `_pos_len` = rows of `pos` that describe the current sequence
        self._rot_stale = False
        self._pos_len = 0
        if pos_mode not in ("append", "shrink"):
            raise ValueError("pos_mode must be in ['append', 'shrink']")
        self.pos_mode = pos_mode          # StreamingArgs.pos_mode of the session this cache belongs to (streaming_args.py:1-10)
        self.pos_upper = 0                # append mode: host-side upper bound of the largest position handed out
        self._status_pinned = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._status_event = None
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
            self.ensure_rotated()
            self._pending_start = self.seq_len
            self._seen_tokens += t
        start = self._pending_start
        if start + t > self.capacity:
            raise RuntimeError(f"StreamingCache capacity {self.capacity} exceeded ({start + t})")
        if layer_idx == self.layers - 1:
            self.seq_len = start + t
        return start

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        """Append raw K / V `[1, H, T, D]` for one layer; returns the `[1, H, S, D]` views of the RAW keys and the values, like
        the reference (:30-74; rotation happens in its attention, language_forward.py:197).  For callers that drive the cache
        the HF way.  The rotated copy of the new rows is written by the same append kernel the patched attention uses
        (`svlm_qkv_rope_append`, append-only form) when the positions of those rows are already known (`compute_positions` /
        `append_positions` ran for this forward); otherwise the cache is marked stale and `ensure_rotated` re-derives K_rot
        as soon as they are."""
        t = key_states.shape[-2]
        start = self.layer_start(layer_idx, t)
        if self._pos_len >= start + t and not self._rot_stale:
            k = key_states.reshape(self.kv_heads, t, self.head_dim).transpose(0, 1).reshape(t, -1).to(torch.bfloat16).contiguous()
            v = value_states.reshape(self.kv_heads, t, self.head_dim).transpose(0, 1).reshape(t, -1).to(torch.bfloat16).contiguous()
            _abi.qkv_rope_append(self.view, self.cur, layer_idx, None, k, v, t, 0, None, start, self.pos, self.cos_table, self.sin_table,
                                 self.section)
        else:
            self.arena[self.cur, K_RAW, layer_idx, :, start:start + t].copy_(key_states.reshape(self.kv_heads, t, self.head_dim))
            self.arena[self.cur, V_PLANE, layer_idx, :, start:start + t].copy_(value_states.reshape(self.kv_heads, t, self.head_dim))
            self._rot_stale = True
        n = start + t
        return (self.arena[self.cur, K_RAW, layer_idx, :, :n].unsqueeze(0), self.arena[self.cur, V_PLANE, layer_idx, :, :n].unsqueeze(0))

    def ensure_rotated(self):
        """Honour `_rot_stale`: re-derive K_rot of every live row from K_raw at the positions in `self.pos` (one identity
        compaction launch into the other buffer).  Raises if rows were replaced and no positions have been computed for them
        since -- attention over keys rotated for a different sequence would be silently wrong."""
        if not self._rot_stale:
            return
        n = self.seq_len
        if n == 0:
            self._rot_stale = False
            return
        if self._pos_len < n:
            raise RuntimeError(f"StreamingCache: K_raw rows were written through update()/key_cache[i] = ... and no positions are known for "
                               f"them ({self._pos_len} < {n} rows): call compute_positions(ids) (shrink mode) before the next attention")
        self.compact([(0, 0, n)])

    # ------------------------------------------------------------------ B200 path
    def set_video_grids(self, grids):
        """grids: iterable of (t, h, w) or an `[n, 3]` tensor, consumed from index 0 by `compute_positions`, as the reference's
        ever-growing `streaming_args.video_grid_thw` is (inference.py:415).  A CUDA tensor is used as is (no host round trip)."""
        if isinstance(grids, torch.Tensor):
            g = grids.reshape(-1, 3).to(device=self.device, dtype=torch.int32).contiguous()
        else:
            g = torch.as_tensor(np.asarray(grids, dtype=np.int32).reshape(-1, 3)).to(self.device)
        self.grids = g
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

    def compute_positions(self, ids, all_text: bool = False) -> int:
        """Contiguous mRoPE positions of `ids` into `self.pos[:, :n]`; `self.next_pos` = max + 1.  `all_text`
        (`StreamingArgs.all_text`, model_forward.py:6-28 `get_1d_rope_index`): every token advances all three axes by one."""
        n = self.upload_ids(ids)
        _abi.mrope_positions(self.ids_dev, n, self.grids, self.n_grids, self.pos, self.next_pos, self.status,
                             self.spatial_merge, self.t_step, -1 if all_text else TOKEN_IDS["<|video_pad|>"])
        self._note_status()
        self._pos_len = n
        if self._rot_stale and n >= self.seq_len:
            self.ensure_rotated()
        return n

    def _note_status(self):
        """Start the read-back of the position kernel's status word; `check_status` (called where the host synchronises
        anyway, after the first token of a turn) raises on a video run that disagrees with its grid."""
        self._status_pinned.copy_(self.status, non_blocking=True)
        if self._status_event is None:
            self._status_event = torch.cuda.Event()
        self._status_event.record()

    def check_status(self):
        if self._status_event is None:
            return
        self._status_event.synchronize()
        st = int(self._status_pinned[0])
        if st & 1:
            raise ValueError("a run of <|video_pad|> tokens does not match its (t, h, w) grid (t * h * w / merge^2 tokens expected)")
        if st & 2:
            raise ValueError("more video runs in the ids than grids in video_grid_thw")

    def reserve_positions(self, n: int):
        """append mode: `n` more positions are about to be handed out; the rotary tables must cover them."""
        self.pos_upper += int(n)
        if self.pos_upper >= self.cos_table.shape[0]:
            raise RuntimeError(f"append-mode positions would exceed the rotary table ({self.cos_table.shape[0]} rows): "
                               "construct the cache / engine with a larger max_position")

    # -- append position mode (reference: model_forward.py:112-153 + language_forward.py:184-188) -----------------
    def gather_positions(self, runs):
        """Retained tokens KEEP their positions in append mode: move them with the same runs as the K/V rows, so that the
        compaction kernel re-derives exactly the K_rot the token was appended with."""
        n = sum(r[2] for r in runs)
        self.pos[:, :n] = torch.cat([self.pos[:, s_:s_ + l] for s_, _, l in runs], dim=1)

    def append_positions(self, ids, start: int, grid, all_text: bool = False) -> int:
        """Positions of the uncached ids `ids[start:]` only: `get_rope_index` over the new ids with THIS chunk's grid, shifted by
        last_cache_position + 1 (= `next_pos`, which the decode steps keep advancing); nothing is recomputed for the cache.
        `grid`: (t, h, w), an `[n, 3]` tensor of the grids of this call, or None (no video among the new ids)."""
        n = self.upload_ids(ids)
        return self._append_positions_uploaded(start, n, grid, all_text)

    def append_new_positions(self, new_ids, grid, all_text: bool = False) -> int:
        """Same for callers that only hold the NEW ids (the reference's append-mode `model_forward` sees the ids sliced to the
        uncached suffix, model_forward.py:112-153): they are placed behind the cached rows."""
        start = self.seq_len
        new = new_ids.reshape(-1) if isinstance(new_ids, torch.Tensor) else torch.from_numpy(as_host_ids(new_ids))
        n = start + new.numel()
        if n > self.capacity:
            raise RuntimeError(f"sequence of {n} tokens exceeds the arena capacity {self.capacity}")
        self.ids_dev[start:n].copy_(new.to(torch.int64), non_blocking=True)
        return self._append_positions_uploaded(start, n, grid, all_text)

    def _append_positions_uploaded(self, start, n, grid, all_text):
        n_new = n - start
        offset = self.next_pos.clone()
        if start == 0:
            offset.zero_()
        if grid is not None:
            self.set_video_grids(grid if isinstance(grid, torch.Tensor) else [grid])
        if not hasattr(self, "_pos_rel"):
            self._pos_rel = torch.zeros_like(self.pos)
        _abi.mrope_positions(self.ids_dev[start:], n_new, self.grids, self.n_grids, self._pos_rel, self.next_pos, self.status,
                             self.spatial_merge, self.t_step, -1 if all_text else TOKEN_IDS["<|video_pad|>"])
        self._note_status()
        self.pos[:, start:n] = self._pos_rel[:, :n_new] + offset
        self.next_pos += offset
        self._pos_len = n
        return n

    def record_decode_positions(self, first_row: int):
        """append mode: the decode kernels rotate at `next_pos` without recording it; the position of every cached row is needed
        again at the next compaction.  Rows first_row .. seq_len-1 hold consecutive text positions ending at next_pos - 1."""
        n_dec = self.seq_len - first_row
        if n_dec > 0:
            self.pos[:, first_row:self.seq_len] = (self.next_pos - n_dec) + torch.arange(n_dec, dtype=torch.int32, device=self.device)
            self._pos_len = max(self._pos_len, self.seq_len)

    def compact(self, runs, new_ids=None):
        """Gather `runs` = [(src, dst, len)] from the live buffer into the other one, re-rotating K.  Shrink mode: at the
        contiguous positions of `new_ids` (the pruned sequence; omitted = positions already in `self.pos`).  Append mode: every
        retained token keeps its position (gathered with the same runs); `new_ids` only refreshes the device copy of the ids."""
        runs = [tuple(int(x) for x in r) for r in runs]
        if len(runs) > _abi.MAX_RUNS:
            raise RuntimeError(f"gather has {len(runs)} runs; SVLM_MAX_RUNS is {_abi.MAX_RUNS} (the eviction policy produces <= 8)")
        n = sum(r[2] for r in runs)
        if self.pos_mode == "append":
            if new_ids is not None:
                self.upload_ids(new_ids)
            if not (len(runs) == 1 and runs[0][:2] == (0, 0)):
                self.gather_positions(runs)
            self._pos_len = n
        elif new_ids is not None:
            stale, self._rot_stale = self._rot_stale, False      # this compaction re-derives K_rot; no separate refresh first
            self.compute_positions(new_ids)
            self._rot_stale = stale
        if self._pos_len < n:
            raise RuntimeError(f"compact: positions are known for {self._pos_len} rows, the gather keeps {n}")
        if runs:
            _abi.kv_compact_rope(self.view, self.cur, runs, self.pos, self.cos_table, self.sin_table, self.section)
        self.cur ^= 1
        self.seq_len = n
        self._rot_stale = False
        self.compactions += 1

    def planes(self, layer):
        self.ensure_rotated()
        b = self.arena[self.cur]
        n = self.seq_len
        return b[K_RAW, layer, :, :n], b[K_ROT, layer, :, :n], b[V_PLANE, layer, :, :n]
