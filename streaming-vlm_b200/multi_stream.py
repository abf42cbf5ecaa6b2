"""Several independent streams on ONE GPU ("one or more streams per GPU", BASELINE config 4).

The reference runs one stream per process (`eval/livesports3kcc/distributed_generate_streaming.py:53-62`): every
stream pays its own sweep of the 15 GB of weights for every decoded token.  Streams are independent, but the weights
are not: here the streams keep their own session state, K/V arena and eviction schedule (one `StreamingEngine` each,
sharing one copy of the weights), the eviction / ViT / chunk prefill of each stream run exactly as in the single-stream
engine, and the DECODE steps of all streams advance together:

* dense projections see a `[B, hidden]` activation matrix -- one weight sweep for B tokens (cuBLAS: a plain library
  GEMM; the M = 1 GEMV kernels of the single-stream path do not apply);
* `svlm_qkv_rope_append_streams` appends token b to stream b's arena at that stream's row / text position;
* `svlm_attn_decode_streams` is one split-KV launch over all streams (grid = splits x kv heads x streams): B x the bytes
  of a single-stream decode attention per launch, which is what lifts it from latency-bound to HBM-bound;
* each step is one CUDA-graph replay; per-stream kv length, next position, token and an `active` flag live in
  device memory, so streams that finished their turn (EOS / token budget) simply stop advancing.

Results per stream are those of the single-stream engine up to the GEMM-vs-GEMV rounding of the decode projections
(`tests/test_gpu_multistream.py`).
"""
from __future__ import annotations

import contextlib
import os
import time

import torch
import torch.nn.functional as F

from . import _abi
from .config import Geometry
from .engine import IM_END, StreamingEngine


class MultiStreamEngine:
    def __init__(self, weights: dict, geo: Geometry, n_streams: int, capacity: int = 8192, device="cuda", share_weights_from=None,
                 **engine_kwargs):
        if not 1 <= n_streams <= _abi.MAX_STREAMS:
            raise ValueError(f"1..{_abi.MAX_STREAMS} streams per engine, got {n_streams}")
        first = StreamingEngine(weights, geo, capacity, device, share_weights_from=share_weights_from, **engine_kwargs)
        self.engines = [first] + [StreamingEngine(None, geo, capacity, device, share_weights_from=first, **engine_kwargs)
                                  for _ in range(n_streams - 1)]
        self.geo, self.device, self.n = geo, first.device, n_streams
        dev, B = self.device, n_streams
        # device-side counters of all streams, one tensor each; the engines' own scalars become views of them
        self.kv_len = torch.zeros(B, dtype=torch.int32, device=dev)
        self.kv_after = torch.zeros(B, dtype=torch.int32, device=dev)
        self.next_pos = torch.zeros(B, dtype=torch.int32, device=dev)
        self.tok = torch.zeros(B, dtype=torch.int64, device=dev)
        self.active = torch.zeros(B, dtype=torch.int32, device=dev)
        for b, e in enumerate(self.engines):
            e.kv_len_dev, e.kv_after_dev = self.kv_len[b:b + 1], self.kv_after[b:b + 1]
            e.tok_dev, e.cache.next_pos = self.tok[b:b + 1], self.next_pos[b:b + 1]
        self._tok_pinned = torch.zeros(B, dtype=torch.int64).pin_memory()
        self._active_pinned = torch.zeros(B, dtype=torch.int32).pin_memory()
        self._tok_ring = [torch.zeros(B, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._tok_events = [torch.cuda.Event() for _ in range(2)]
        self._undo = torch.zeros(B, dtype=torch.int32, device=dev)
        self._keys = torch.zeros(B, dtype=torch.int64, device=dev)          # packed (logit, ~id) maxima of the lm_head launch, zero between steps
        # pinned staging buffers of the look-ahead loop's mask uploads, used round-robin: a buffer is rewritten two iterations later
        # at the earliest, after the host has waited for an event recorded behind its copy
        self._mask_ring = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(8)]
        self._mask_next = 0
        self.lookahead = os.environ.get("SVLM_MULTI_LOOKAHEAD", "1") != "0"
        g = geo
        self._q_rot = torch.zeros(B, g.num_heads, g.head_dim, dtype=torch.bfloat16, device=dev)
        self._o = torch.zeros(B, g.num_heads, g.head_dim, dtype=torch.bfloat16, device=dev)
        # one wave of CTAs (splits x kv heads x streams <= 148 SMs) and >= ~3 key tiles per CTA: swept in
        # benchmarks/streams_attn_sweep.py (B = 8: 4 splits, 22.9 us per layer for 95 MB = 0.63 of the HBM peak)
        self.max_splits = max(1, min((capacity + 127) // 128, 148 // (g.num_kv_heads * B)))
        self.ws = torch.zeros(_abi.attn_workspace_bytes(B, g.num_heads, g.head_dim, self.max_splits) // 4 + 16, dtype=torch.float32,
                              device=dev)
        self._graphs, self._slots = {}, {}
        self._pf_graphs, self._pf_seen, self._pf_calls = {}, {}, 0
        self._pf_ids = torch.zeros(B * 1024, dtype=torch.int64, device=dev)
        self.batch_prefill = True      # ViT and chunk prefill of all streams in one pass (one weight sweep); False = per stream
        self.concurrent_prefill = os.environ.get("SVLM_CONCURRENT_PREFILL", "1") != "0"   # per-stream append + attention on side streams
        self._pf_streams = [torch.cuda.Stream(device=dev) for _ in range(B)]
        self.own_gemm = os.environ.get("SVLM_BATCHED_DECODE_CUBLAS", "0") != "1"      # decode projections: svlm_gemm_tokens (default) / cuBLAS
        self.use_graph = first.use_graph
        self.launches = 0
        self.timing = []

    # ------------------------------------------------------------------ batched decode step
    def _decode_body(self, slots, splits):
        e0, g, B = self.engines[0], self.geo, self.n
        hq, hkv, d = g.num_heads, g.num_kv_heads, g.head_dim
        c0 = e0.cache
        before = e0.launches
        x = torch.empty(B, g.hidden, dtype=torch.bfloat16, device=self.device)
        _abi.embed_tokens(self.tok, B, e0.embed, None, 0, x, e0._pf_status)
        self.launches += 1
        if not self.own_gemm:
            return self._decode_body_cublas(x, slots, splits, before)
        # Five launches per layer, every projection ONE launch of the weight-streaming swap-AB GEMM (svlm_gemm_tokens: weight rows on
        # the tensor-core M dimension, the B tokens on N, split-K over a cluster) with everything around it fused:
        #   RMSNorm -> q/k/v projection -> bias -> mRoPE -> append to each stream's arena | split-KV attention over all streams |
        #   o_proj + residual | RMSNorm -> gate/up projection -> SwiGLU | down_proj + residual;   final norm -> lm_head -> argmax.
        # The two row-parallel launches leave per-tile sums of squares of the rows they wrote (`stats`), from which the NEXT launch's
        # fused RMSNorm gets its row statistics with one short read (layer 0 computes them from the embedding rows in the kernel).
        dev, bf = self.device, torch.bfloat16
        act = torch.empty(B, g.mlp, dtype=bf, device=dev)
        st_attn = torch.empty(-(-g.hidden // 128), B, dtype=torch.float32, device=dev)
        st_mlp = torch.empty_like(st_attn)
        for li, L in enumerate(e0.layers):
            _abi.gemm_tokens_qkv_append(c0.view, slots, li, x, L["wqkv"], L["bqkv"], L["ln1"], g.rms_eps, hq, self._q_rot, c0.cos_table,
                                        c0.sin_table, stats_in=st_mlp if li else None)
            _abi.attn_decode_streams(c0.view, slots, li, self._q_rot, self._o, hq, c0.capacity, e0.scale, splits,
                                     self.ws if splits > 1 else None, early_kv=e0._early_kv)
            _abi.gemm_tokens(self._o.view(B, -1), L["wo"], x, residual=x, epilogue=_abi.EPI_RESIDUAL, stats_out=st_attn)
            _abi.gemm_tokens(x, L["wgu"], act, epilogue=_abi.EPI_SWIGLU, norm_weight=L["ln2"], eps=g.rms_eps, stats_in=st_attn)
            _abi.gemm_tokens(act, L["wd"], x, residual=x, epilogue=_abi.EPI_RESIDUAL, stats_out=st_mlp)
        # final norm + lm_head + suppress mask + argmax in the projection's epilogue (no logits tensor), then ONE small launch that hands
        # every active stream its token and advances the device-side counters
        _abi.gemm_tokens_lm_head_argmax(x, e0.lm_head, self._keys, norm_weight=e0.norm, eps=g.rms_eps, stats_in=st_mlp,
                                        suppress_u8=e0.suppress_u8)
        _abi.streams_advance(self._keys, self.tok, self.active, self.kv_len, self.kv_after, self.next_pos)
        merged_in_cluster = splits in (2, 4, 8)                # the batched attention merges these split counts inside clusters: no merge launch
        self.launches += g.num_layers * (5 + (1 if splits > 1 and not merged_in_cluster else 0)) + 2

    def _advance(self, nxt):
        act = self.active
        self.tok.copy_(torch.where(act.bool(), nxt, self.tok))
        self.kv_len.add_(act)
        self.kv_after.copy_(self.kv_len + 1)
        self.next_pos.add_(act)

    def _decode_body_cublas(self, x, slots, splits, before):
        """The same step with the projections in cuBLAS (SVLM_BATCHED_DECODE_CUBLAS=1: the A/B baseline of the own kernels)."""
        e0, g, B = self.engines[0], self.geo, self.n
        hq, hkv, d = g.num_heads, g.num_kv_heads, g.head_dim
        c0 = e0.cache
        pending = None
        for li, L in enumerate(e0.layers):
            x, h = e0._add_norm(x, pending, L["ln1"], g.rms_eps)
            qkv = F.linear(h, L["wqkv"], L["bqkv"])
            q, k, v = qkv[:, :hq * d], qkv[:, hq * d:(hq + hkv) * d], qkv[:, (hq + hkv) * d:]
            _abi.qkv_rope_append_streams(c0.view, slots, li, q, k, v, hq, self._q_rot, c0.cos_table, c0.sin_table, c0.section)
            _abi.attn_decode_streams(c0.view, slots, li, self._q_rot, self._o, hq, c0.capacity, e0.scale, splits,
                                     self.ws if splits > 1 else None, early_kv=e0._early_kv)
            x, h = e0._add_norm(x, F.linear(self._o.view(B, -1), L["wo"]), L["ln2"], g.rms_eps)
            pending = F.linear(e0._swiglu(F.linear(h, L["wgu"]), g.mlp), L["wd"])
        logits = e0._logits(x, pending)                                   # [B, vocab] fp32, suppressed ids at -inf
        self._advance(torch.argmax(logits, dim=-1))
        self.launches += (e0.launches - before) + g.num_layers * (2 + (1 if splits > 1 and splits not in (2, 4, 8) else 0))
        e0.launches = before

    def decode_step(self, done=None):
        """Every ACTIVE stream consumes its token in `self.tok`, appends its K/V and leaves its next greedy token there.
        `done[b]`: stream b finished its turn (its counters do not advance).  Like the single-stream engine, a stream that
        would append beyond its arena raises instead of dropping the row."""
        caches = [e.cache for e in self.engines]
        for b, c in enumerate(caches):
            if (done is None or not done[b]) and c.seq_len + 1 > c.capacity:
                raise RuntimeError(f"stream {b}: arena capacity {c.capacity} exceeded during decode")
        n_tiles = max(c.seq_len for c in caches) // 128 + 1
        splits = max(1, min(self.max_splits, -(-n_tiles // 3)))
        key = (tuple(c.cur for c in caches), splits)
        slots = self._slots.get(key[0])
        if slots is None:
            slots = self._slots[key[0]] = _abi.make_slots(caches, list(self.kv_len.split(1)), list(self.next_pos.split(1)),
                                                          list(self.kv_after.split(1)))
        if not self.use_graph:
            self._decode_body(slots, splits)
            return
        entry = self._graphs.get(key)
        if entry is None:
            saved = [t.clone() for t in (self.kv_len, self.kv_after, self.next_pos, self.tok)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._decode_body(slots, splits)                           # warm-up outside capture (cuBLAS workspaces)
            torch.cuda.current_stream().wait_stream(side)
            for t, s in zip((self.kv_len, self.kv_after, self.next_pos, self.tok), saved):
                t.copy_(s)
            torch.cuda.synchronize()
            before = self.launches
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                self._decode_body(slots, splits)
            entry = self._graphs[key] = (gr, self.launches - before)
            self.launches = before
        entry[0].replay()
        self.launches += entry[1]

    # ------------------------------------------------------------------ batched ViT + chunk prefill
    def _prefill_body(self, ids, metas, video, last):
        """Chunk prefill of all streams in ONE pass over the weights: the dense projections see the concatenated new tokens of
        every stream; append and attention run per stream on its own arena (rows / lengths from device memory).
        metas[b] = (row offset, T_b, video-token runs, splits); video = [sum(video tokens), hidden] in stream order."""
        e0, g = self.engines[0], self.geo
        hq, hkv, d = g.num_heads, g.num_kv_heads, g.head_dim
        total = ids.numel()
        before = e0.launches
        x = F.embedding(ids, e0.embed)
        vo = 0
        for off, t, runs, _ in metas:
            for a, n in runs:
                x[off + a:off + a + n] = video[vo:vo + n]
                vo += n
        pending = None
        for li, L in enumerate(e0.layers):
            x, h = e0._add_norm(x, pending, L["ln1"], g.rms_eps)
            qkv = F.linear(h, L["wqkv"], L["bqkv"])
            q, k, v = qkv[:, :hq * d], qkv[:, hq * d:(hq + hkv) * d], qkv[:, (hq + hkv) * d:]
            q_rot = torch.empty(total, hq, d, dtype=torch.bfloat16, device=self.device)
            o = torch.empty(total, hq, d, dtype=torch.bfloat16, device=self.device)
            # Append + attention of the streams are independent of each other (own arena, own rows of q / o): they fork onto one side
            # stream per stream and run CONCURRENTLY, each as ONE launch over the whole KV range (no split, no merge kernel) -- a
            # single stream's 68 CTAs leave half of the SMs idle and need a 2-way split + merge to fill them; two streams side by side
            # fill them with full-length chains (the regime in which this kernel reaches 0.55 of the tensor peak).  Inside a CUDA-graph
            # capture the fork / join becomes graph edges.
            main = torch.cuda.current_stream()
            concurrent = self.concurrent_prefill and len(self.engines) > 1
            for b, (e, (off, t, _, splits)) in enumerate(zip(self.engines, metas)):
                c, sl = e.cache, slice(off, off + t)
                if concurrent:
                    side = self._pf_streams[b]
                    side.wait_stream(main)
                    splits = 1
                ctx = torch.cuda.stream(side) if concurrent else contextlib.nullcontext()
                with ctx:
                    _abi.qkv_rope_append(c.view, c.cur, li, q[sl], k[sl], v[sl], t, hq, q_rot[sl], 0, c.pos, c.cos_table, c.sin_table, c.section,
                                         start_dev=e.kv_len_dev)
                    _abi.attn_causal(c.view, c.cur, li, q_rot[sl], o[sl], t, hq, c.capacity, e.scale, splits, e.ws if splits > 1 else None,
                                     kv_len_dev=e.kv_after_dev)
                self.launches += 2 + (1 if splits > 1 else 0)
            if concurrent:
                for b in range(len(self.engines)):
                    main.wait_stream(self._pf_streams[b])
            x, h = e0._add_norm(x, F.linear(o.view(total, -1), L["wo"]), L["ln2"], g.rms_eps)
            pending = F.linear(e0._swiglu(F.linear(h, L["wgu"]), g.mlp), L["wd"])
        logits = e0._logits(x[last], pending[last])           # `last`: row of the last new token of every stream
        self.tok.copy_(torch.argmax(logits, dim=-1))
        self.launches += e0.launches - before
        e0.launches = before

    def _vit_and_prefill_all(self, ids_all, pixels, grid, video=None):
        """Batched ViT (the chunks of all streams are the temporal slices of one (B*t, h, w) grid: windows and full attention
        are per slice, so this is exactly B independent towers) followed by the batched chunk prefill."""
        import numpy as np
        from .engine import VIDEO_PAD
        engines, g = self.engines, self.geo
        e0 = engines[0]
        if video is None:
            video = e0.vision_forward(torch.cat(pixels, dim=0), (grid[0] * self.n, grid[1], grid[2]))
        metas, off, pos = [], 0, 0
        for e, ids in zip(engines, ids_all):
            c = e.cache
            new_ids = np.asarray(ids[c.seq_len:])
            t, start = len(new_ids), c.seq_len
            if start + t > c.capacity:
                raise RuntimeError(f"sequence of {start + t} tokens exceeds the arena capacity {c.capacity}")
            where = np.flatnonzero(new_ids == VIDEO_PAD)
            brk = np.flatnonzero(np.diff(where) != 1) + 1
            runs = tuple((int(where[a]), int(b - a)) for a, b in zip(np.concatenate([[0], brk]), np.concatenate([brk, [where.size]]))) if where.size else ()
            splits = e.prefill_splits or (1 if (-(-t // (128 // g.group))) * g.num_kv_heads >= 74 or start + t < 2048 or t > 512 else 2)
            metas.append((off, t, runs, splits))
            self._pf_ids[off:off + t].copy_(c.ids_dev[start:start + t])
            e.kv_len_dev.fill_(start)
            e.kv_after_dev.fill_(start + t)
            off += t
            pos += int(where.size)
        if pos != video.shape[0]:
            raise ValueError(f"Video features and video tokens do not match: tokens: {pos}, features {video.shape[0]}")
        ids = self._pf_ids[:off]
        key = (tuple(metas), tuple(e.cache.cur for e in engines), video.data_ptr())
        entry = self._pf_graphs.get(key) if self.use_graph else None
        if entry is None:
            last = torch.tensor([o_ + t - 1 for o_, t, _, _ in metas], dtype=torch.long, device=self.device)
            # A layout (new-token counts of all streams) is captured when it shows up during start-up or for the second time: a
            # stream that ends one turn early produces a one-off layout, and capturing that would cost far more than running it
            # eagerly once.
            self._pf_calls += 1
            self._pf_seen[key] = self._pf_seen.get(key, 0) + 1
            if not self.use_graph or (self._pf_calls > 4 and self._pf_seen[key] < 2):
                self._prefill_body(ids, metas, video, last)
            else:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._prefill_body(ids, metas, video, last)            # warm-up outside capture (appends the same rows again)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                before = self.launches
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._prefill_body(ids, metas, video, last)
                entry = self._pf_graphs[key] = (gr, self.launches - before, last)
                self.launches = before
        if entry is not None:
            entry[0].replay()
            self.launches += entry[1]
        for e, (_, t, _, _) in zip(engines, metas):
            e.cache.seq_len += t
            e.cache._seen_tokens += t
            e.kv_len_dev.fill_(e.cache.seq_len)
            e.kv_after_dev.fill_(e.cache.seq_len + 1)

    # ------------------------------------------------------------------ decode loops
    def _decode_loop_plain(self, ids_all, seqs, done, argmaxes, forced):
        """One batched decode step per token, the host reads every token before the next step (teacher forcing needs that)."""
        B, engines = self.n, self.engines
        n_fwd = 1
        while True:
            self._tok_pinned.copy_(self.tok, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            changed = False
            for b in range(B):
                if done[b]:
                    continue
                tok = int(self._tok_pinned[b])
                argmaxes[b].append(tok)
                f = None if forced is None else forced[b]
                if f is not None and len(argmaxes[b]) <= len(f) and f[len(argmaxes[b]) - 1] != tok:
                    tok = int(f[len(argmaxes[b]) - 1])
                    engines[b].tok_dev.fill_(tok)
                seqs[b].append(tok)
                if tok == IM_END or len(seqs[b]) - len(ids_all[b]) >= engines[b].max_new_tokens:
                    done[b] = changed = True
            if all(done):
                break
            if changed:
                for b in range(B):
                    self._active_pinned[b] = 0 if done[b] else 1
                self.active.copy_(self._active_pinned, non_blocking=True)
            self.decode_step(done)
            n_fwd += 1
            for b, e in enumerate(engines):
                if not done[b]:
                    e.cache.seq_len += 1
                    e.cache._seen_tokens += 1
        return n_fwd

    def _upload_mask(self, dst, values):
        buf = self._mask_ring[self._mask_next]
        self._mask_next = (self._mask_next + 1) % len(self._mask_ring)
        for b, v in enumerate(values):
            buf[b] = 1 if v else 0
        dst.copy_(buf, non_blocking=True)

    def _decode_loop_lookahead(self, ids_all, seqs, done, argmaxes):
        """The same loop with one step of look-ahead (as in StreamingEngine.generate): the next batched step is enqueued BEFORE the
        host has seen the current tokens, so the GPU never idles on the device->host round trip and the host-side bookkeeping.  A
        stream whose token fills its budget is switched off ahead of the speculated step; a stream whose token turns out to be
        <|im_end|> is rolled back afterwards (its speculated K/V row lies beyond the live length, only the counters moved)."""
        B, engines = self.n, self.engines
        n_fwd = 1
        ring = self._tok_ring
        step = 0
        while True:
            slot = step & 1
            ring[slot].copy_(self.tok, non_blocking=True)
            self._tok_events[slot].record()
            # streams for which the token being read is the last one of their budget do not take part in the next step
            last = [(not done[b]) and len(seqs[b]) + 1 - len(ids_all[b]) >= engines[b].max_new_tokens for b in range(B)]
            spec = [not done[b] and not last[b] for b in range(B)]
            speculated = any(spec) and all(engines[b].cache.seq_len + 1 <= engines[b].cache.capacity for b in range(B) if spec[b])
            if speculated:
                if any(last):
                    self._upload_mask(self.active, spec)
                self.decode_step([not x for x in spec])
            self._tok_events[slot].synchronize()
            ended = []
            for b in range(B):
                if done[b]:
                    continue
                tok = int(ring[slot][b])
                argmaxes[b].append(tok)
                seqs[b].append(tok)
                if tok == IM_END or last[b]:
                    done[b] = True
                    if speculated and spec[b]:
                        ended.append(b)
                elif speculated:
                    engines[b].cache.seq_len += 1
                    engines[b].cache._seen_tokens += 1
            if speculated:
                n_fwd += 1
            if ended:                                       # undo the speculated step of the streams that had just ended
                self._upload_mask(self._undo, [b in ended for b in range(B)])
                self.kv_len.sub_(self._undo)
                self.next_pos.sub_(self._undo)
                self.kv_after.copy_(self.kv_len + 1)
                self._upload_mask(self.active, [not d for d in done])
            if all(done):
                break
            if not speculated:
                self._upload_mask(self.active, [not d for d in done])
                self.decode_step(done)                      # raises the capacity error
                n_fwd += 1
                for b, e in enumerate(engines):
                    if not done[b]:
                        e.cache.seq_len += 1
                        e.cache._seen_tokens += 1
            step += 1
        return n_fwd

    # ------------------------------------------------------------------ one chunk of every stream
    def step(self, prompts, pixels, grid, previous_text="", forced=None, sync_timing=True):
        """One chunk for each stream (`prompts[b]`, `pixels[b]`; same grid).  `forced[b]`: teacher-forced tokens of stream b
        (see StreamingEngine.generate).  Returns the per-stream result dicts of StreamingEngine.step."""
        B, engines = self.n, self.engines
        if sync_timing:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        # The vision tower needs nothing but the new frames: its one batched graph replay is enqueued FIRST, and the host-side
        # eviction planning of all streams (~1.2 ms each, the only part of a round that does not batch) runs while the GPU is busy
        # with it.  The compaction launches queue up behind it in stream order.
        video_all = None
        same_shape = all(p is not None for p in pixels) and len({tuple(p.shape) for p in pixels}) == 1
        if self.batch_prefill and same_shape:
            video_all = engines[0].vision_forward(torch.cat(pixels, dim=0), (grid[0] * self.n, grid[1], grid[2]))
        ids_all, runs_all = [], []
        for b, e in enumerate(engines):
            ids, runs = e.begin_step(prompts[b], grid, previous_text)
            ids_all.append(ids)
            runs_all.append(runs)
        if sync_timing:
            torch.cuda.synchronize()
        t_pkv = time.perf_counter() - t0
        t0 = time.perf_counter()
        batched = (self.batch_prefill and all(p is not None for p in pixels) and len({tuple(p.shape) for p in pixels}) == 1
                   and all(len(ids_all[b]) - e.cache.seq_len >= 2 for b, e in enumerate(engines))
                   and sum(len(ids_all[b]) - e.cache.seq_len for b, e in enumerate(engines)) <= self._pf_ids.numel())
        if batched:
            self._vit_and_prefill_all(ids_all, pixels, grid, video_all)
        else:
            per = None if video_all is None else video_all.shape[0] // B
            for b, e in enumerate(engines):
                c = e.cache
                if video_all is not None:
                    video = video_all[b * per:(b + 1) * per]
                else:
                    video = e.vision_forward(pixels[b], grid) if pixels[b] is not None else None
                e.prefill(ids_all[b][c.seq_len:], video)            # leaves the stream's first token in e.tok_dev (= self.tok[b])
                e.kv_len_dev.fill_(c.seq_len)
                e.kv_after_dev.fill_(c.seq_len + 1)
        seqs = [list(ids) for ids in ids_all]
        done = [False] * B
        argmaxes = [[] for _ in range(B)]
        self.active.fill_(1)
        n_fwd = 1
        if forced is None and self.use_graph and self.lookahead:
            n_fwd = self._decode_loop_lookahead(ids_all, seqs, done, argmaxes)
        else:
            n_fwd = self._decode_loop_plain(ids_all, seqs, done, argmaxes, forced)
        if sync_timing:
            torch.cuda.synchronize()
        t_gen = time.perf_counter() - t0
        self.last_argmax = argmaxes
        self.timing.append({"PKV": t_pkv, "GEN": t_gen, "n_forward": n_fwd})
        return [e.finish_step(ids_all[b], seqs[b], runs_all[b], t_pkv / B, t_gen / B, n_fwd) for b, e in enumerate(engines)]

    def total_launches(self):
        return self.launches + sum(e.launches for e in self.engines)
