"""Per-stream session runner: evict -> ViT -> chunk prefill -> greedy decode, one stream per object.

This is the B200-native counterpart of the reference's `streaming_inference` chunk loop
(`inference/inference.py:309-517`) together with the patched forwards it drives
(`qwen2_5/model_forward.py:30-176`, `qwen2_5/language_forward.py:143-305, 395-516`,
`qwen2_5/vision_forward.py:6-101`) and the greedy branch of `_sample`
(`generate/streaming_generate_qwen.py:53-110`).  Differences in execution, not in results:

* the K/V cache is a fixed arena (`StreamingCache`); eviction is ONE compaction launch per chunk,
  positions are recomputed on the device once per chunk (the reference: host Python on every forward);
* K is rotated once when it is appended and once per eviction (the reference re-rotates the whole
  cache in every layer of every forward);
* GQA is resolved inside the attention kernel (no `repeat_kv` copies); decode attention is split-KV;
* lm_head runs on the last position only (the reference computes all and samples `[:, -1]`);
* each decode step is one CUDA-graph replay (kv length and next position live in device memory).

The dense projections of the ViT and of the chunk prefill go through cuBLAS (`torch.nn.functional.linear`);
every other op on the path -- frame ingest, eviction, positions, appends, attention, the patch merger and the
whole decode forward (weight-streaming GEMVs with fused norm / bias / rotation / append / SwiGLU / residual /
argmax) -- is a kernel of `libsvlm_b200.so`.  Section timers mirror the reference's PKV / GEN
(`inference.py:312-329, 422-452`).
"""
from __future__ import annotations

import math
import os
import time
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import _abi
from .config import Geometry
from .inference.generate.streaming_cache import StreamingCache
from .inference.inference import (DEFAULT_TEXT_ROUND, DEFAULT_TEXT_SINK, DEFAULT_TEXT_SLIDING_WINDOW, DEFAULT_WINDOW_SIZE,
                                  MAX_TOKEN_PER_DURATION, plan_past_kv)
from .utils.get_qwen_range import TOKEN_IDS

IM_END = TOKEN_IDS["<|im_end|>"]
LF = TOKEN_IDS["\n"]
VIDEO_PAD = TOKEN_IDS["<|video_pad|>"]


def _window_index(grid, window_size, merge, patch):
    """HF `get_window_index` for one (t, h, w) grid -> (order of merge units, cu_seqlens in ViT tokens)."""
    t, h, w = grid
    ws = window_size // merge // patch
    gh, gw = h // merge, w // merge
    idx = np.arange(t * gh * gw).reshape(t, gh, gw)
    ph, pw = ws - gh % ws, ws - gw % ws
    nh, nw = (gh + ph) // ws, (gw + pw) // ws
    pad = np.full((t, gh + ph, gw + pw), -100, dtype=np.int64)
    pad[:, :gh, :gw] = idx
    pad = pad.reshape(t, nh, ws, nw, ws).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, ws, ws)
    lens = (pad != -100).sum(axis=(2, 3)).reshape(-1)
    flat = pad.reshape(-1)
    order = flat[flat != -100]
    cu = np.concatenate([[0], np.cumsum(lens) * merge * merge])
    cu = cu[np.concatenate([[True], np.diff(cu) != 0])]           # unique_consecutive (vision_forward.py:71)
    return order, cu.astype(np.int32)


class _VitPlan:
    """Everything about one (t, h, w) grid that does not depend on pixel values (cached per grid)."""

    def __init__(self, grid, geo: Geometry, device):
        t, h, w = grid
        m = geo.spatial_merge_size
        unit = m * m
        self.n = t * h * w
        order, cu_win = _window_index(grid, geo.window_size, m, geo.patch_size)
        hd = geo.vit_hidden // geo.vit_heads
        dim = hd // 2
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
        hp = torch.arange(h)[:, None].expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w)[None, :].expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        pid = torch.stack([hp, wp], dim=-1).repeat(t, 1)
        table = torch.outer(torch.arange(max(h, w), dtype=torch.float), inv)
        emb = table[pid].flatten(1)
        order_t = torch.from_numpy(order)
        emb = emb.reshape(self.n // unit, unit, -1)[order_t].reshape(self.n, -1)
        emb = torch.cat((emb, emb), dim=-1)
        self.cos = emb.cos().contiguous().to(device)
        self.sin = emb.sin().contiguous().to(device)
        self.order = order_t.to(device)
        self.out_rows = order_t.to(torch.int32).to(device)       # merged row m lands at row order[m] (argsort un-permute)
        self.cu_window = torch.from_numpy(cu_win).to(device)
        self.cu_full = torch.tensor([i * h * w for i in range(t + 1)], dtype=torch.int32, device=device)


@dataclass(frozen=True)
class GenConfig:
    """Token selection of one turn: the reference generates with `do_sample=True, temperature=0.9, repetition_penalty=1.05`
    (`inference.py:41-42, 426-451`; logits processors + multinomial at `generate/streaming_generate_qwen.py:73-99`); greedy
    (`do_sample=False`, penalty 1, temperature 1) is the parity mode.  All of it runs inside the lm_head kernel
    (`svlm_lm_head_sample`); a frozen value because CUDA graphs are keyed on it."""
    do_sample: bool = False
    repetition_penalty: float = 1.0
    temperature: float = 1.0
    seed: int = None        # None: every turn draws from the engine's own stream (base seed + turn counter); an int pins the turn


GREEDY = GenConfig()
REFERENCE_SAMPLING = GenConfig(do_sample=True, repetition_penalty=1.05, temperature=0.9)
PREFILL_BUCKET = 8          # chunk-prefill graphs are captured for token counts rounded up to a multiple of this


def decode_split_count(seq_len: int, num_kv_heads: int, limit: int = 1 << 30) -> int:
    """Split count of the single-stream decode attention for a cache of `seq_len` rows: ~2 key tiles per CTA keeps the serial
    TMA -> UMMA -> softmax chain short, at most one wave of CTAs (148 SMs / kv heads); the merge kernel takes up to 48 partials in one
    L2 round trip (swept in benchmarks/microbench.py)."""
    n_tiles = (seq_len + 128) // 128
    return max(1, min(limit, 148 // num_kv_heads, -(-n_tiles // 2)))


class _GraphCache(OrderedDict):
    """LRU-bounded map of CUDA graphs: the keys carry run-time values (token-count bucket, split count, live buffer,
    generation config), so the number of captured graphs must not grow with the length of the stream."""

    def __init__(self, limit):
        super().__init__()
        self.limit = limit
        self.evictions = 0

    def get(self, key):
        v = super().get(key)
        if v is not None:
            self.move_to_end(key)
        return v

    def put(self, key, value):
        self[key] = value
        self.move_to_end(key)
        while len(self) > self.limit:
            self.popitem(last=False)
            self.evictions += 1
        return value


class StreamingEngine:
    def __init__(self, weights: dict, geo: Geometry, capacity: int = 8192, device="cuda", window_size=DEFAULT_WINDOW_SIZE,
                 text_round=DEFAULT_TEXT_ROUND, text_sink=DEFAULT_TEXT_SINK, text_sliding_window=DEFAULT_TEXT_SLIDING_WINDOW,
                 max_new_tokens=MAX_TOKEN_PER_DURATION, suppress=None, assistant_start_bias=3, assistant_end_bias=2,
                 use_graph=True, decode_splits=None, prefill_splits=None, share_weights_from=None, pos_mode="shrink",
                 max_position=None, tp_group=None, gen: GenConfig = GREEDY, rebind=None, max_graphs: int = 24, cache=None,
                 prefill_bucket: int = PREFILL_BUCKET):
        if not torch.cuda.is_available():
            raise RuntimeError("StreamingEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        _abi.lib()
        # head-sharded tensor parallelism (parallel.shard_llm_weights): `geo` and `weights` are this rank's shard, `tp_group` the
        # NCCL group whose all-reduce sums the o_proj / down_proj partial products
        # NCCL group whose ranks hold the other shards.  The two row-parallel projections of a layer are reduced by our own kernels over
        # symmetric memory (parallel.TpContext): no NCCL call on the data path, so the steps stay CUDA-graph capturable (replaying
        # graphs that contained NCCL all-reduces hung on this pool in round 1, which had forced this mode to run eagerly).
        self.tp_group = tp_group
        self.tp = None
        if tp_group is not None:
            from .parallel import TpContext
            self.tp = TpContext(tp_group, geo.hidden, max_rows=1024, device=device)
        self.geo, self.device = geo, torch.device(device)
        self.window_size, self.text_round = window_size, text_round
        self.text_sink, self.text_sliding_window = text_sink, text_sliding_window
        self.max_new_tokens = max_new_tokens
        self.asb, self.aeb = assistant_start_bias, assistant_end_bias
        self.use_graph = use_graph
        self.suppress = None if suppress is None else suppress.to(self.device)
        # StreamingArgs.pos_mode (streaming_args.py:1-10): "shrink" re-indexes the retained tokens contiguously after every
        # eviction (the reference's default); "append" keeps every token's position for ever ("grows indefinitely"), so the
        # rotary tables must cover the whole stream: `max_position` rows (default 131072 = ~40 min of 448x448 video at 2 fps)
        if pos_mode not in ("append", "shrink"):
            raise ValueError("pos_mode must be in ['append', 'shrink']")
        self.pos_mode = pos_mode
        if pos_mode == "append" and max_position is None:
            max_position = 1 << 17
        if cache is not None:            # an arena the caller owns (the drop-in's `past_key_values`)
            capacity = cache.capacity
            self.cache = cache
        else:
            self.cache = StreamingCache(geo, capacity, self.device, max_position=max_position, pos_mode=pos_mode)
        self.gen = gen
        self.prefill_bucket = max(1, int(prefill_bucket))
        self.video_pixel_budget = (None, None)      # (min_pixels, max_pixels) per frame for `ingest`; None = keep sizes that fit the patch grid
        self.scale = 1.0 / math.sqrt(geo.head_dim)
        self.decode_splits = decode_splits or max(1, min((capacity + 127) // 128, 296 // geo.num_kv_heads))   # two CTAs per SM
        self.prefill_splits = prefill_splits
        if share_weights_from is not None:      # another stream of the same model on this GPU: one copy of the weights
            for name in ("w_patch", "vit", "vit_mlp_padded", "merger", "embed", "layers", "norm", "lm_head"):
                setattr(self, name, getattr(share_weights_from, name))
        else:
            self._prepare_weights(weights, rebind)
        g = geo
        dev = self.device
        self.ws = torch.zeros(max(_abi.attn_workspace_bytes(1, g.num_heads, g.head_dim, self.decode_splits),
                                  _abi.attn_workspace_bytes(512, g.num_heads, g.head_dim, 4)) // 4 + 16, dtype=torch.float32, device=dev)
        self.kv_len_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.kv_after_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.tok_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.tok_pinned = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._tok_ring = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._tok_events = [torch.cuda.Event() for _ in range(2)]
        self._graphs = _GraphCache(max_graphs)
        self._dec_qkv = torch.zeros((g.num_heads + 2 * g.num_kv_heads) * g.head_dim, dtype=torch.bfloat16, device=dev)
        self._dec_q = torch.zeros(1, g.num_heads, g.head_dim, dtype=torch.bfloat16, device=dev)
        # side stream of the decode step: L2 prefetch of the o_proj weights and the head of the gate/up weights while the
        # (latency-bound, HBM-idle) attention runs; SVLM_DECODE_PREFETCH_MB = MiB of gate/up to pull (0 disables the branch)
        self._side = torch.cuda.Stream(device=dev)
        self._prefetch_gu_bytes = int(float(os.environ.get("SVLM_DECODE_PREFETCH_MB", "16")) * (1 << 20))
        self._prefetch = os.environ.get("SVLM_DECODE_PREFETCH", "1") != "0"
        self._early_kv = os.environ.get("SVLM_ATTN_EARLY_KV", "1") != "0"
        self._dec_o = torch.zeros(1, g.num_heads, g.head_dim, dtype=torch.bfloat16, device=dev)
        self._dec_act = torch.zeros(g.mlp, dtype=torch.bfloat16, device=dev)
        self._dec_x = torch.zeros(g.hidden, dtype=torch.bfloat16, device=dev)
        self._dec_part = torch.zeros(g.hidden, dtype=torch.bfloat16, device=dev)
        self._argmax_key = torch.zeros(1, dtype=torch.int64, device=dev)
        self._pf_ids = torch.zeros(1024, dtype=torch.int64, device=dev)
        self._pf_last = torch.zeros(1, dtype=torch.int64, device=dev)          # row of the last real token in a padded prefill
        self._pf_video = torch.zeros(1024, g.hidden, dtype=torch.bfloat16, device=dev)   # static home of the chunk's video embeddings
        self._pf_status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._pf_status_pinned = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._scratch_cnt = torch.zeros(3, dtype=torch.int32, device=dev)      # counters the prefill's lm_head launch may advance
        self.seen = torch.zeros(g.vocab, dtype=torch.uint8, device=dev)        # ids in the sequence so far (repetition penalty)
        self.rng_dev = torch.zeros(4, dtype=torch.int32, device=dev)           # Philox state: seed lo, seed hi, step (tokens drawn this turn)
        self.base_seed, self._turns = 0, 0
        self._rng_host = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.suppress_u8 = None if self.suppress is None else self.suppress.to(torch.uint8).contiguous()
        self._vit_plans = {}
        self._vit_graphs = _GraphCache(8)
        self._pf_graphs = _GraphCache(max_graphs)
        # session state (reference: inference.py:297-303)
        self.prev_ids = None
        self.history, self.clips, self.pixels, self.grids = [], [], [], []
        self.i = 0
        self.timing = []
        self.launches = 0

    # ------------------------------------------------------------------ weights
    def _prepare_weights(self, w, rebind=None):
        """`w`: HF parameter names -> tensors (consumed).  Projections that the kernels read as ONE matrix ([q;k;v], [gate;up],
        zero-padded ViT MLP) are fused here.  `rebind(name, view)`: called for every parameter that was fused, with a view of
        the fused tensor that holds the same values -- `from_hf_model` uses it to re-point the live HF module's parameters, so
        model and engine share ONE copy of the weights (everything that is not fused is aliased as is)."""
        g = self.geo
        dev = self.device
        mv = lambda t: t.to(dev).contiguous()
        rb = rebind or (lambda name, view: None)

        def fuse_rows(names, pad_to=None):
            parts = [w.pop(n) for n in names]
            rows = [p_.shape[0] for p_ in parts]
            if pad_to is not None:
                parts = [F.pad(p_, (0, 0, 0, pad_to - p_.shape[0])) if p_.dim() == 2 else F.pad(p_, (0, pad_to - p_.shape[0])) for p_ in parts]
            fused = mv(torch.cat(parts, 0))
            off = 0
            for n, r, p_ in zip(names, rows, parts):
                rb(n, fused[off:off + r])
                off += p_.shape[0]
            return fused

        v = "model.visual."
        pe = w.pop(v + "patch_embed.proj.weight")
        self.w_patch = mv(pe.reshape(pe.shape[0], -1))
        self.vit = []
        # The ViT MLP width (3420) makes the down projection's rows 6840 B -- not 16-byte aligned, which pushes cuBLAS onto a
        # legacy mma.sync kernel (58 us vs ~12 us).  Zero-padding the width to a multiple of 64 changes no result:
        # silu(0) * 0 = 0 feeds zero columns of the down projection.
        mp = -(-g.vit_mlp // 64) * 64
        self.vit_mlp_padded = mp
        for li in range(g.vit_depth):
            b = f"{v}blocks.{li}."
            wd = mv(F.pad(w.pop(b + "mlp.down_proj.weight"), (0, mp - g.vit_mlp)))
            rb(b + "mlp.down_proj.weight", wd[:, :g.vit_mlp])
            self.vit.append(dict(
                n1=mv(w.pop(b + "norm1.weight")), n2=mv(w.pop(b + "norm2.weight")),
                wqkv=mv(w.pop(b + "attn.qkv.weight")), bqkv=mv(w.pop(b + "attn.qkv.bias")),
                wproj=mv(w.pop(b + "attn.proj.weight")), bproj=mv(w.pop(b + "attn.proj.bias")),
                wgu=fuse_rows([b + "mlp.gate_proj.weight", b + "mlp.up_proj.weight"], mp),
                bgu=fuse_rows([b + "mlp.gate_proj.bias", b + "mlp.up_proj.bias"], mp),
                wd=wd, bd=mv(w.pop(b + "mlp.down_proj.bias"))))
        self.merger = dict(ln=mv(w.pop(v + "merger.ln_q.weight")), w0=mv(w.pop(v + "merger.mlp.0.weight")),
                           b0=mv(w.pop(v + "merger.mlp.0.bias")), w2=mv(w.pop(v + "merger.mlp.2.weight")),
                           b2=mv(w.pop(v + "merger.mlp.2.bias")))
        lm = "model.language_model."
        self.embed = mv(w.pop(lm + "embed_tokens.weight"))
        self.layers = []
        for li in range(g.num_layers):
            b = f"{lm}layers.{li}."
            a = b + "self_attn."
            self.layers.append(dict(
                ln1=mv(w.pop(b + "input_layernorm.weight")), ln2=mv(w.pop(b + "post_attention_layernorm.weight")),
                wqkv=fuse_rows([a + "q_proj.weight", a + "k_proj.weight", a + "v_proj.weight"]),
                bqkv=fuse_rows([a + "q_proj.bias", a + "k_proj.bias", a + "v_proj.bias"]),
                wo=mv(w.pop(a + "o_proj.weight")),
                wgu=fuse_rows([b + "mlp.gate_proj.weight", b + "mlp.up_proj.weight"]),
                wd=mv(w.pop(b + "mlp.down_proj.weight"))))
        self.norm = mv(w.pop(lm + "norm.weight"))
        self.lm_head = mv(w.pop("lm_head.weight")) if "lm_head.weight" in w else self.embed      # tied embeddings

    @classmethod
    def from_hf_model(cls, model, capacity: int = 8192, **kwargs):
        """Engine over the parameters of a live (patched) HF Qwen2.5-VL model, no second copy of the weights: unfused tensors
        are aliased, fused ones ([q;k;v], [gate;up], padded ViT MLP) replace the module's own parameters by views of the fused
        tensor.  This is what `convert_qwen2_5_to_streaming` + `generate` run on (reference `qwen2_5/patch_model.py:18-38`)."""
        from .config import from_hf_config
        geo = from_hf_config(model.config)
        if geo.vit_kind != "qwen2_5":
            raise RuntimeError("the session-runner fast path covers the Qwen2.5-VL tower; Qwen2-VL runs the patched eager forwards")
        params = dict(model.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("StreamingEngine.from_hf_model needs the model on a CUDA device (there is no CPU fallback)")
        bad = [n for n, p_ in params.items() if p_.dtype != torch.bfloat16]
        if bad:
            raise RuntimeError(f"the svlm kernels compute in bf16; parameters such as {bad[0]} are {params[bad[0]].dtype}")

        def rebind(name, view):
            params[name].data = view

        w = {n: p_.data for n, p_ in params.items()}
        with torch.no_grad():
            return cls(w, geo, capacity=capacity, device=dev, rebind=rebind, **kwargs)

    # ------------------------------------------------------------------ drop-in plumbing
    def adopt_cache(self, cache: StreamingCache):
        """Run on another arena of the same geometry (`past_key_values` handed to the patched `generate`).  CUDA graphs bake the
        arena's addresses in, so the captured ones are dropped."""
        g = self.geo
        if (cache.layers, cache.kv_heads, cache.head_dim) != (g.num_layers, g.num_kv_heads, g.head_dim):
            raise RuntimeError("StreamingCache geometry does not match the model")
        if cache.arena.device != self.embed.device:
            raise RuntimeError(f"StreamingCache lives on {cache.arena.device}, the model on {self.embed.device}")
        self.cache = cache
        self._graphs.clear()
        self._pf_graphs.clear()
        self.decode_splits = max(1, min((cache.capacity + 127) // 128, 296 // g.num_kv_heads))
        need = max(_abi.attn_workspace_bytes(1, g.num_heads, g.head_dim, self.decode_splits),
                   _abi.attn_workspace_bytes(512, g.num_heads, g.head_dim, 4)) // 4 + 16
        if self.ws.numel() < need:
            self.ws = torch.zeros(need, dtype=torch.float32, device=self.device)

    def set_suppress(self, mask):
        """Ids that are never emitted (HF `suppress_tokens`; the synthetic streams use it to keep a random-init model inside the
        token grammar).  Updated in place when possible: the mask's address is part of the captured graphs."""
        mask = mask.to(self.device)
        if self.suppress is not None and self.suppress.shape == mask.shape:
            if self.suppress.data_ptr() != mask.data_ptr():
                self.suppress.copy_(mask)
                self.suppress_u8.copy_(mask)
            return
        self.suppress = mask.clone()
        self.suppress_u8 = mask.to(torch.uint8).contiguous()
        self._graphs.clear()
        self._pf_graphs.clear()

    # ------------------------------------------------------------------ small helpers
    def _mark(self, name):
        """Optional phase timeline (CUDA events on the current stream), enabled by `self.phase_marks = []`."""
        marks = getattr(self, "phase_marks", None)
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e, time.perf_counter()))

    def _row_parallel(self, which, a, w, x):
        """Tensor-parallel rank: x + sum over ranks of (a @ w.T).  The GEMM writes this rank's partial products straight into the
        symmetric buffer; `svlm_tp_allreduce_rows` reduces them over NVLink peer loads and adds the residual."""
        t = a.shape[0]
        part = self.tp.rows_view(which, t, w.shape[0])
        torch.mm(a, w.t(), out=part)
        out = torch.empty_like(x)
        _abi.tp_allreduce_rows(self.tp.ref, which, x, out, t, w.shape[0])
        self.tp.all_reduces += 1
        self.launches += 1
        return out

    def _rmsnorm(self, x, w, eps=1e-6):
        out = torch.empty_like(x)
        _abi.rmsnorm(x, w, out, eps)
        self.launches += 1
        return out

    def _add_norm(self, x, pending, w, eps=1e-6):
        """(x + pending, RMSNorm(x + pending) * w) in one kernel; pending=None -> (x, RMSNorm(x) * w)."""
        h = torch.empty_like(x)
        if pending is None:
            _abi.add_rmsnorm(x, None, w, None, h, eps)
            self.launches += 1
            return x, h
        xs = torch.empty_like(x)
        _abi.add_rmsnorm(x, pending, w, xs, h, eps)
        self.launches += 1
        return xs, h

    def _swiglu(self, gu, m):
        out = torch.empty(gu.shape[0], m, dtype=gu.dtype, device=gu.device)
        _abi.swiglu(gu, out)
        self.launches += 1
        return out

    # ------------------------------------------------------------------ ViT (vision_forward.py:57-101)
    def vision_forward(self, pixels: torch.Tensor, grid, taps=None):
        """ViT on one chunk.  With `use_graph` the whole tower (~420 launches, static shapes per grid) is one CUDA-graph
        replay: the eager tower is host-launch-bound (4.6 ms of Python for 2 ms of GPU work at the 7B geometry)."""
        g = self.geo
        grid = tuple(int(x) for x in grid)
        plan = self._vit_plans.get(grid)
        if plan is None:
            plan = self._vit_plans[grid] = _VitPlan(grid, g, self.device)
        if not self.use_graph or taps is not None:
            return self._vision_body(pixels, plan, taps)
        key = (grid, pixels.dtype)
        entry = self._vit_graphs.get(key)
        if entry is None:
            px_in = torch.zeros_like(pixels)
            px_in.copy_(pixels)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._vision_body(px_in, plan, None)            # warm-up outside capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            before = self.launches
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                out = self._vision_body(px_in, plan, None)
            entry = self._vit_graphs[key] = (gr, px_in, out, self.launches - before)
            self.launches = before
        gr, px_in, out, n_launch = entry
        px_in.copy_(pixels, non_blocking=True)
        gr.replay()
        self.launches += n_launch
        return out

    def _vision_body(self, pixels, plan, taps=None):
        g = self.geo
        n, unit = plan.n, g.spatial_merge_size ** 2
        heads, hd = g.vit_heads, g.vit_hidden // g.vit_heads
        x = F.linear(pixels.to(torch.bfloat16), self.w_patch)
        x = x.reshape(n // unit, unit, -1)[plan.order].reshape(n, -1)
        pending = None                                      # sub-layer output not yet added to the residual stream
        for li, b in enumerate(self.vit):
            cu = plan.cu_full if li in g.fullatt_block_indexes else plan.cu_window
            x, h = self._add_norm(x, pending, b["n1"])
            qkv = F.linear(h, b["wqkv"], b["bqkv"])
            _abi.vit_rope(qkv, plan.cos, plan.sin, n, heads, hd)
            o = torch.empty(n, heads, hd, dtype=torch.bfloat16, device=self.device)
            _abi.vit_attn(qkv, cu, o, n, heads, hd, hd ** -0.5)
            self.launches += 2
            x, h = self._add_norm(x, F.linear(o.view(n, -1), b["wproj"], b["bproj"]), b["n2"])
            gu = F.linear(h, b["wgu"], b["bgu"])
            pending = F.linear(self._swiglu(gu, self.vit_mlp_padded), b["wd"], b["bd"])
        _, h = self._add_norm(x, pending, self.merger["ln"])
        h = h.view(n // unit, -1)
        y = torch.empty(n // unit, self.merger["w0"].shape[0], dtype=torch.bfloat16, device=self.device)
        _abi.gemm_bf16(h, self.merger["w0"], self.merger["b0"], y, act=1)
        out = torch.empty(n // unit, g.hidden, dtype=torch.bfloat16, device=self.device)
        _abi.gemm_bf16(y, self.merger["w2"], self.merger["b2"], out, act=0, out_rows=plan.out_rows)
        self.launches += 2
        return out

    # ------------------------------------------------------------------ LLM layers (language_forward.py:143-305)
    def _layer(self, x, pending, li, t, start, kv_len, splits, taps=None, dev_start=False):
        """One decoder layer of the chunk prefill.  `pending` is the previous layer's MLP output, added to the residual
        stream inside the fused add+RMSNorm kernel; returns (x, pending) in the same convention."""
        g, c, L = self.geo, self.cache, self.layers[li]
        hq, hkv, d = g.num_heads, g.num_kv_heads, g.head_dim
        x, h = self._add_norm(x, pending, L["ln1"], g.rms_eps)
        qkv = F.linear(h, L["wqkv"], L["bqkv"])
        q, k, v = qkv[:, :hq * d], qkv[:, hq * d:(hq + hkv) * d], qkv[:, (hq + hkv) * d:]
        q_rot = torch.empty(t, hq, d, dtype=torch.bfloat16, device=self.device)
        if dev_start:        # graph-captured prefill: arena row and kv length come from device memory
            _abi.qkv_rope_append(c.view, c.cur, li, q, k, v, t, hq, q_rot, 0, c.pos, c.cos_table, c.sin_table, c.section,
                                 start_dev=self.kv_len_dev)
        else:
            _abi.qkv_rope_append(c.view, c.cur, li, q, k, v, t, hq, q_rot, start, c.pos, c.cos_table, c.sin_table, c.section)
        o = torch.empty(t, hq, d, dtype=torch.bfloat16, device=self.device)
        _abi.attn_causal(c.view, c.cur, li, q_rot, o, t, hq, kv_len, self.scale, splits, self.ws if splits > 1 else None,
                         kv_len_dev=self.kv_after_dev if dev_start else None)
        self.launches += 2 + (1 if splits > 1 else 0)
        if taps is not None:
            taps.setdefault("attn_out", []).append(o.view(t, -1))
            taps.setdefault("q_rot", []).append(q_rot)
        if self.tp is None:
            x, h = self._add_norm(x, F.linear(o.view(t, -1), L["wo"]), L["ln2"], g.rms_eps)
            gu = F.linear(h, L["wgu"])
            return x, F.linear(self._swiglu(gu, g.mlp), L["wd"])
        x = self._row_parallel(0, o.view(t, -1), L["wo"], x)
        _, h = self._add_norm(x, None, L["ln2"], g.rms_eps)
        gu = F.linear(h, L["wgu"])
        return self._row_parallel(1, self._swiglu(gu, g.mlp), L["wd"], x), None

    def _logits(self, x_last, pending_last=None):
        _, h = self._add_norm(x_last, pending_last, self.norm, self.geo.rms_eps)
        logits = F.linear(h, self.lm_head).float()
        if self.suppress is not None:
            logits = logits.masked_fill(self.suppress, float("-inf"))
        return logits

    def _gen_key(self, gen):
        # the seed is NOT part of the key: it lives in device memory (`rng_dev`), a new seed needs no new graph
        return (gen.do_sample, float(gen.repetition_penalty), float(gen.temperature))

    def _select_token(self, x_last, pending_last, gen, counters=None):
        """Final norm + lm_head + token selection (`gen`) in ONE kernel pair: the chosen id lands in `tok_dev`.  `counters`:
        the (kv_len, kv_after, next_pos) device scalars the launch advances (decode) -- scratch for the prefill."""
        g = self.geo
        x = x_last.reshape(-1)
        if pending_last is not None:
            x = x + pending_last.reshape(-1)
        kv_len, kv_after, next_pos = counters if counters is not None else self._scratch_cnt.split(1)
        _abi.lm_head_sample(self.lm_head, x, self.norm, g.rms_eps, self.suppress_u8, self._argmax_key, self.tok_dev, kv_len, kv_after, next_pos,
                            seen=self.seen, repetition_penalty=gen.repetition_penalty, temperature=gen.temperature,
                            do_sample=gen.do_sample, rng_state=self.rng_dev)
        self.launches += 2

    def prefill(self, new_ids, video_embeds, taps=None, gen: GenConfig = None):
        """Forward over the uncached ids (already uploaded as the tail of `cache.ids_dev`, positions in `cache.pos`); leaves the
        first token of the turn in `tok_dev` (selected by `gen` inside the lm_head kernel).
        With `use_graph`, one CUDA graph per (token-count bucket, splits, live buffer, generation config): the arena row, the kv
        length, the number of real tokens and the layout of the video placeholders are read from device memory, so the same
        graph serves every chunk whose new-token count falls into the bucket; rows beyond the real count are padding (causal
        attention: they come after every real row, their K/V rows lie beyond the live length and are overwritten by decode)."""
        c, g = self.cache, self.geo
        gen = gen or self.gen
        c.ensure_rotated()
        t, start = len(new_ids), c.seq_len
        if start + t > c.capacity:
            raise RuntimeError(f"sequence of {start + t} tokens exceeds the arena capacity {c.capacity}")
        n_video = 0
        if video_embeds is not None:
            n_video = video_embeds.shape[0]
            n_pad = int(np.count_nonzero(np.asarray(new_ids) == VIDEO_PAD))
            if n_pad != n_video:
                raise ValueError(f"Video features and video tokens do not match: tokens: {n_pad}, features {n_video}")
        t_pad = min(-(-t // self.prefill_bucket) * self.prefill_bucket, c.capacity - start)
        graphed = self.use_graph and taps is None and t >= 2 and t_pad <= self._pf_ids.numel() and n_video <= self._pf_video.shape[0]
        if not graphed:
            t_pad = t
        splits = self.prefill_splits or (1 if (-(-t_pad // (128 // g.group))) * g.num_kv_heads >= 74 or start + t_pad < 2048 or t_pad > 512 else 2)
        if not graphed:
            self._prefill_body(c.ids_dev[start:start + t], t, t, start, video_embeds, n_video, splits, False, taps, gen)
        else:
            self._pf_ids[:t].copy_(c.ids_dev[start:start + t])
            if t_pad > t:
                self._pf_ids[t:t_pad].zero_()
            self._pf_last.fill_(t - 1)
            if n_video:
                self._pf_video[:n_video].copy_(video_embeds)
            self.kv_len_dev.fill_(start)
            self.kv_after_dev.fill_(start + t_pad)
            key = (t_pad, splits, c.cur, n_video > 0, self._gen_key(gen))
            n_cap = self._pf_video.shape[0] if n_video else 0     # baked into the graph: the static buffer's size, not this chunk's count
            entry = self._pf_graphs.get(key)
            if entry is None:
                saved = (self.seen.clone(), self.rng_dev.clone())
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._prefill_body(self._pf_ids[:t_pad], t_pad, t, start, self._pf_video, n_cap, splits, True, None, gen)   # warm-up
                torch.cuda.current_stream().wait_stream(side)
                self.seen.copy_(saved[0]); self.rng_dev.copy_(saved[1])
                torch.cuda.synchronize()
                before = self.launches
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._prefill_body(self._pf_ids[:t_pad], t_pad, t, start, self._pf_video, n_cap, splits, True, None, gen)
                entry = self._pf_graphs.put(key, (gr, self.launches - before))
                self.launches = before
            gr, n_launch = entry
            gr.replay()
            self.launches += n_launch
        c.seq_len = start + t
        c._seen_tokens += t

    def _prefill_body(self, ids, t_pad, t, start, video_embeds, n_video, splits, dev_start, taps, gen):
        g = self.geo
        x = torch.empty(t_pad, g.hidden, dtype=torch.bfloat16, device=self.device)
        _abi.embed_tokens(ids, t_pad, self.embed, video_embeds if n_video else None, n_video, x, self._pf_status, VIDEO_PAD)
        self.launches += 1
        pending = None
        for li in range(g.num_layers):
            x, pending = self._layer(x, pending, li, t_pad, start, self.cache.capacity if dev_start else start + t_pad, splits, taps, dev_start)
        if dev_start:           # padded prefill: the last REAL row is named by a device scalar
            x_last, p_last = x.index_select(0, self._pf_last), None if pending is None else pending.index_select(0, self._pf_last)
        else:
            x_last, p_last = x[t - 1:t], None if pending is None else pending[t - 1:t]
        if taps is not None:
            taps["logits"] = self._logits(x_last.contiguous(), None if p_last is None else p_last.contiguous())[0]
        self._select_token(x_last, p_last, gen)

    # ------------------------------------------------------------------ decode (one token per CUDA-graph replay)
    def _decode_body(self, gen: GenConfig = GREEDY):
        """One decode forward, every op a kernel of libsvlm_b200.so except the embedding row lookup: per layer
        norm+QKV GEMV -> rotate/append -> split-KV attention (+merge) -> o_proj GEMV + residual -> norm+gate/up GEMV
        with SwiGLU -> down GEMV + residual; then final norm + lm_head GEMV + argmax + counter update."""
        c, g = self.cache, self.geo
        hq, hkv, d = g.num_heads, g.num_kv_heads, g.head_dim
        x = self._dec_x                                 # embedding row of the token (own kernel: one launch instead of ATen's gather + copy)
        _abi.embed_tokens(self.tok_dev, 1, self.embed, None, 0, x.view(1, -1), self._pf_status, VIDEO_PAD)
        self.launches += 1
        q_rot, o, act = self._dec_q, self._dec_o, self._dec_act
        for li, L in enumerate(self.layers):
            _abi.gemv_qkv_rope_append(c.view, c.cur, li, L["wqkv"], x, L["ln1"], g.rms_eps, L["bqkv"], hq, q_rot, self.kv_len_dev,
                                      c.next_pos, c.cos_table, c.sin_table)
            if self._prefetch:
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    _abi.prefetch_l2(L["wo"])
                    if self._prefetch_gu_bytes:
                        _abi.prefetch_l2(L["wgu"], self._prefetch_gu_bytes)
            # early_kv: the counters were advanced at the end of the previous token and the launch right before this one appends
            # row kv_len - 1 only, so the attention kernel may fetch the older K/V tiles before its grid dependency resolves
            _abi.attn_causal(c.view, c.cur, li, q_rot, o, 1, hq, c.capacity, self.scale, self._splits_now,
                             self.ws if self._splits_now > 1 else None, kv_len_dev=self.kv_after_dev, early_kv=self._early_kv)
            if self._prefetch:
                main.wait_stream(self._side)
            if self.tp is None:
                _abi.gemv_bf16(L["wo"], o.view(-1), x, 1, residual=x)
                _abi.gemv_bf16(L["wgu"], x, act, 2, L["ln2"], g.rms_eps)
                _abi.gemv_bf16(L["wd"], act, x, 1, residual=x)
            else:       # row-parallel projections: GEMV + one-shot all-reduce over NVLink + residual add, one kernel each
                _abi.tp_gemv_allreduce(self.tp.ref, 0, L["wo"], o.view(-1), x, x)
                _abi.gemv_bf16(L["wgu"], x, act, 2, L["ln2"], g.rms_eps)
                _abi.tp_gemv_allreduce(self.tp.ref, 1, L["wd"], act, x, x)
                self.tp.all_reduces += 2
        self.launches += g.num_layers * (5 + (1 if self._splits_now > 1 else 0) + ((2 if self._prefetch_gu_bytes else 1) if self._prefetch else 0))
        _abi.lm_head_sample(self.lm_head, x, self.norm, g.rms_eps, self.suppress_u8, self._argmax_key, self.tok_dev,
                            self.kv_len_dev, self.kv_after_dev, c.next_pos, seen=self.seen, repetition_penalty=gen.repetition_penalty,
                            temperature=gen.temperature, do_sample=gen.do_sample, rng_state=self.rng_dev)
        self.launches += 2

    def decode_step(self, gen: GenConfig = None):
        """Consumes `tok_dev`, appends its K/V, leaves the next token (selected by `gen`) in `tok_dev`."""
        c = self.cache
        gen = gen or self.gen
        if c.seq_len + 1 > c.capacity:
            raise RuntimeError(f"arena capacity {c.capacity} exceeded during decode")
        # split count for the current cache length: ~2 key tiles per CTA keeps the serial TMA->UMMA->softmax chain short, at most
        # one wave of CTAs; the merge kernel takes up to 48 partials in one L2 round trip (swept in benchmarks/microbench.py)
        self._splits_now = decode_split_count(c.seq_len, self.geo.num_kv_heads, self.decode_splits)
        if self.use_graph:
            key = (c.cur, self._splits_now, self._gen_key(gen))
            gr = self._graphs.get(key)
            if gr is None:
                # warm up once outside capture (cuBLAS workspaces, lazy kernel loading), then restore the state it advanced
                saved = (self.kv_len_dev.clone(), c.next_pos.clone(), self.tok_dev.clone(), self.kv_after_dev.clone(), self.seen.clone(),
                         self.rng_dev.clone())
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._decode_body(gen)
                torch.cuda.current_stream().wait_stream(s)
                self.kv_len_dev.copy_(saved[0]); c.next_pos.copy_(saved[1]); self.tok_dev.copy_(saved[2]); self.kv_after_dev.copy_(saved[3])
                self.seen.copy_(saved[4]); self.rng_dev.copy_(saved[5])
                torch.cuda.synchronize()
                before = self.launches
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._decode_body(gen)
                self._graph_launches = self.launches - before
                self.launches = before
                self._graphs.put(key, gr)
            gr.replay()
            self.launches += self._graph_launches
        else:
            self._decode_body(gen)
        c.seq_len += 1
        c._seen_tokens += 1

    def generate(self, ids_full, pixels, grid, taps=None, forced=None, gen: GenConfig = None, max_new_tokens=None, eos=IM_END, video=None):
        """`model.generate` of the reference for one chunk: ViT + prefill + <= max_new_tokens tokens selected by `gen`
        (default: the engine's own config; `GREEDY` is the parity mode, `REFERENCE_SAMPLING` the reference's).
        `forced` (teacher forcing, the reference's --gt_json idea, inference.py:231-239): token j of the chunk is
        taken from this list instead of the selected one; the selection is still recorded in `self.last_argmax`."""
        c = self.cache
        gen = gen or self.gen
        budget = self.max_new_tokens if max_new_tokens is None else int(max_new_tokens)
        input_len = len(ids_full)
        new_ids = ids_full[c.seq_len:]
        if gen.repetition_penalty != 1.0:
            # the penalty applies to every id of the sequence so far (`input_ids` of `_sample`); cache.ids_dev holds them
            _abi.mark_seen(c.ids_dev, input_len, self.seen, clear_first=True)
            self.launches += 1
        if gen.do_sample:
            # Philox key of this turn: the caller's seed, or the engine's own stream (base seed, turn counter); step word = 0
            seed = gen.seed if gen.seed is not None else ((self.base_seed & 0xFFFFFFFF) << 32) | (self._turns & 0xFFFFFFFF)
            self._rng_host[:] = torch.tensor(_abi.rng_words(seed), dtype=torch.int32)
            self.rng_dev.copy_(self._rng_host, non_blocking=True)
            self._turns += 1
        self._mark("vit_start")
        if video is None:           # (`video`: embeddings of this chunk that the caller already has in flight, see `step`)
            video = self.vision_forward(pixels, grid, taps) if pixels is not None else None
        if taps is not None and video is not None:
            taps["video_embeds"] = video
        self._mark("prefill_start")
        self.prefill(new_ids, video, taps, gen)
        self._pf_status_pinned.copy_(self._pf_status, non_blocking=True)
        self._mark("decode_start")
        self.kv_len_dev.fill_(c.seq_len)
        self.kv_after_dev.fill_(c.seq_len + 1)
        seq = list(ids_full)
        n_fwd = 1
        self.last_argmax = []
        if forced is None and self.use_graph:
            # Decode loop with one step of look-ahead: the next decode step is enqueued BEFORE the host has seen the current
            # token, so the GPU never idles on the device->host round trip.  The token copy is stream-ordered ahead of the
            # step that overwrites `tok_dev`.  If the token turns out to be <|im_end|>, the speculated step is rolled back:
            # its K/V row lies beyond the live length (never read), only the counters are restored.
            pinned = self._tok_ring
            speculated = False
            while True:
                slot = len(self.last_argmax) & 1
                pinned[slot].copy_(self.tok_dev, non_blocking=True)
                ev = self._tok_events[slot]
                ev.record()
                will_be_last = len(seq) + 1 - input_len >= budget     # the token being read fills the budget
                if not will_be_last and c.seq_len + 1 <= c.capacity:
                    self.decode_step(gen)
                    speculated = True
                else:
                    speculated = False
                ev.synchronize()
                tok = int(pinned[slot][0])
                if not self.last_argmax:
                    self._check_chunk_status()
                self.last_argmax.append(tok)
                seq.append(tok)
                if tok == eos or will_be_last:
                    if speculated:                      # undo the step that consumed the end token
                        c.seq_len -= 1
                        c._seen_tokens -= 1
                        c.next_pos.sub_(1)
                        if gen.do_sample:
                            self.rng_dev[2:3].sub_(1)
                    break
                if not speculated:
                    self.decode_step(gen)               # raises the capacity error
                n_fwd += 1
            self._mark("decode_end")
            return seq, n_fwd
        while True:
            self.tok_pinned.copy_(self.tok_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            tok = int(self.tok_pinned[0])
            if not self.last_argmax:
                self._check_chunk_status()
            self.last_argmax.append(tok)
            if forced is not None and len(self.last_argmax) <= len(forced) and forced[len(self.last_argmax) - 1] != tok:
                tok = int(forced[len(self.last_argmax) - 1])
                self.tok_dev.fill_(tok)
                if gen.repetition_penalty != 1.0:
                    self.seen[tok] = 1
            seq.append(tok)
            if tok == eos or len(seq) - input_len >= budget:
                break
            self.decode_step(gen)
            n_fwd += 1
        self._mark("decode_end")
        return seq, n_fwd

    def _check_chunk_status(self):
        """Once per chunk, where the host has just synchronised on the first token anyway: the status words of the position
        kernel (video run vs grid) and of the embedding splice (placeholder count, id range).  Their device->host copies were
        enqueued ahead of the token copy, so nothing extra is waited for."""
        self.cache.check_status()
        st = int(self._pf_status_pinned[0])
        if st:
            self._pf_status.zero_()
        if st & 8:
            raise ValueError("Video features and video tokens do not match: more <|video_pad|> tokens than feature rows")
        if st & 16:
            raise ValueError("token id outside the vocabulary in the prompt")

    def load_state(self, state, seed=0):
        """Adopt the bookkeeping produced by `synthetic.fast_forward` and fill the arena with random K_raw / V for the
        rows it implies (K_rot is then derived by the real kernel), i.e. jump to the steady state of a long stream."""
        c = self.cache
        n = int(state["cache_len"])
        if n > c.capacity:
            raise RuntimeError(f"state needs {n} cache rows, arena capacity is {c.capacity}")
        self.prev_ids = np.asarray(state["prev_ids"], dtype=np.int64)
        self.history, self.clips, self.pixels = state["history"], state["clips"], state["pixels"]
        self.grids, self.i = list(state["grids"]), int(state["i"])
        g = torch.Generator(device=self.device).manual_seed(seed)
        for plane in (0, 2):
            c.arena[c.cur, plane, :, :, :n].normal_(0.0, 0.5, generator=g)
        n_vid = int(np.count_nonzero(self.prev_ids[:n] == TOKEN_IDS["<|vision_start|>"]))
        c.set_video_grids(self.grids[:max(n_vid, 1)])
        self._grids_uploaded = None
        c.seq_len = n
        c.compute_positions(self.prev_ids[:n])
        c.compact([(0, 0, n)])
        c.pos_upper = n

    # ------------------------------------------------------------------ session loop (inference.py:309-517)
    def begin_step(self, prompt_ids, grid, previous_text=""):
        """Eviction half of a chunk (`process_past_kv` + position recompute, inference.py:309-420): plans the pruning of the
        previous ids, compacts + re-rotates the arena, uploads ids and positions.  Returns (ids of this chunk, gather runs)."""
        i, c = self.i, self.cache
        runs, plan = None, None
        if self.prev_ids is not None:
            plan = plan_past_kv(self.prev_ids, i, self.text_round, self.window_size, self.history, self.asb, self.aeb,
                                self.clips, self.pixels, self.text_sink, self.text_sliding_window)
            prev = plan.ids.tolist()
            runs = plan.runs(c.seq_len)
        else:
            prev = None
        self.clips.append(None)
        if i == 0:
            self.history = [{"role": "previous text", "content": previous_text},
                            {"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"},
                                                         {"type": "text", "text": "query"}]}]
        else:
            self.history.append({"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}]})
        prompt = [int(t) for t in prompt_ids]
        ids = prompt if prev is None else prev + (prompt if prev[-1] != LF else prompt[1:])
        self.pixels.append(None)
        self.grids.append(tuple(int(x) for x in grid))
        n_vid = int(np.count_nonzero(np.asarray(ids) == TOKEN_IDS["<|vision_start|>"]))
        want = self.grids[:max(n_vid, 1)]
        if getattr(self, "_grids_uploaded", None) != want:
            c.set_video_grids(want)
            self._grids_uploaded = list(want)
        moved = runs is not None and not (len(runs) == 1 and runs[0] == (0, 0, c.seq_len))
        if self.pos_mode == "append":
            if moved:
                c.compact(runs)                 # append mode: the positions travel with the rows (StreamingCache.compact)
                self.launches += 1
            c.reserve_positions(len(ids) - c.seq_len + self.max_new_tokens)
            c.append_positions(ids, c.seq_len, self.grids[-1])
            self._grids_uploaded = None
            self.launches += 1
            return ids, runs
        c.compute_positions(ids)
        self.launches += 1
        if moved:
            c.compact(runs)
            self.launches += 1
        return ids, runs

    def finish_step(self, ids, seq, runs, t_pkv=0.0, t_gen=0.0, n_fwd=0):
        """Bookkeeping after generation (inference.py:457-517): close the turn, remember the ids, extend the history."""
        input_len = len(ids)
        if self.pos_mode == "append":
            self.cache.record_decode_positions(input_len)
        if seq[-1] != IM_END:
            seq.append(IM_END)
        new_tokens = seq[input_len:]
        self.prev_ids = np.asarray(seq, dtype=np.int64)
        self.history.append({"role": "assistant", "content": " ".join(str(t) for t in new_tokens) + " ..."})
        self.timing.append({"PKV": t_pkv, "GEN": t_gen, "n_forward": n_fwd, "new_tokens": len(new_tokens)})
        self.i += 1
        return {"new_tokens": new_tokens, "input_len": input_len, "ids": seq, "runs": runs, "cache_len": self.cache.seq_len}

    def ingest(self, frames_thwc: torch.Tensor):
        """uint8 RGB frames [T, H, W, 3] in HBM -> (bf16 patch rows for the ViT, grid): the pixel half of the reference's
        `processor(text=..., videos=clip)` call (inference.py:390-395) for a clip already at its target size."""
        g = self.geo
        factor = g.patch_size * g.spatial_merge_size
        t, h, w, _ = frames_thwc.shape
        lo, hi = self.video_pixel_budget
        if h % factor or w % factor or (hi is not None and h * w > hi) or (lo is not None and h * w < lo):
            # livecc_utils' _spatial_resize_video (video_process_patch.py:133-153): smart_resize target, anti-aliased bicubic
            from .utils.frames import smart_resize
            h2, w2 = smart_resize(h, w, factor, lo if lo is not None else factor * factor, hi if hi is not None else max(h * w, factor * factor))
            frames_thwc = _abi.frames_resize_bicubic(frames_thwc.contiguous(), h2, w2)
            self.launches += 2
        self.launches += 1
        return _abi.frames_to_patches(frames_thwc, patch=g.patch_size, temporal_patch=g.temporal_patch_size, merge=g.spatial_merge_size)

    def step(self, prompt_ids, pixels, grid, previous_text="", taps=None, sync_timing=True, forced=None, frames=None):
        """One chunk.  Either `pixels` (patch rows as the HF processor would emit them) or `frames` (uint8 [T, H, W, 3] on
        the device, run through `ingest`) carries the video; `grid` is ignored when `frames` is given."""
        if sync_timing:
            torch.cuda.synchronize()
        self._mark("step_start")
        if frames is not None:
            pixels, grid = self.ingest(frames)
        # The vision tower needs nothing but the new frames: with nothing to tap and no section timing asked for, its graph replay
        # is enqueued FIRST and the host-side eviction planning below (~1 ms) runs while the GPU is busy with it.
        video = None
        if pixels is not None and taps is None and not sync_timing:
            video = self.vision_forward(pixels, grid)
        t0 = time.perf_counter()
        ids, runs = self.begin_step(prompt_ids, grid, previous_text)
        if sync_timing:
            torch.cuda.synchronize()
        t_pkv = time.perf_counter() - t0

        t0 = time.perf_counter()
        seq, n_fwd = self.generate(ids, pixels, grid, taps, forced, video=video)
        if sync_timing:
            torch.cuda.synchronize()
        t_gen = time.perf_counter() - t0
        return self.finish_step(ids, seq, runs, t_pkv, t_gen, n_fwd)
