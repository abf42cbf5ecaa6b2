"""Stream sharding across GPUs: independent video streams, one process per GPU, no data-path collective.

The reference shards inference the same way (`eval/livesports3kcc/distributed_generate_streaming.py:53-62`:
`device_map=f'cuda:{device_id}'`, work items `idxs[device_id::num_workers]`; `eval/model_compete/generate.sh:57-80`).
`torch.distributed` is used only for the timing barrier and to combine per-rank counters
(units summed, time = max over ranks), never for K/V or activations.
"""
from __future__ import annotations

import torch


def assign_streams(n_streams: int, rank: int, world: int):
    """Round-robin partition: stream s runs on rank s % world."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_streams, world))


def aggregate(units_local: float, seconds_local: float, device=None):
    """(total units over all ranks, max seconds over ranks).  Works with any initialised backend (nccl / gloo)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(units_local), float(seconds_local)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())


# ------------------------------------------------------------------------------------------------------------------
# Head-sharded tensor parallelism of the language model (BASELINE configs[4], SURVEY section 8(e) "optional second mode"; the
# reference has no such mode).  Rank r owns q heads [r*Hq/W, (r+1)*Hq/W), kv heads [r*Hkv/W, ...) and a 1/W slice of the MLP
# width: q/k/v/gate/up are column-parallel (rows of the nn.Linear weight), o_proj / down_proj row-parallel (columns), so the
# K/V arena, the compaction, the RoPE and the attention kernels run unchanged on fewer heads, and a layer needs exactly two
# all-reduces (after o_proj and after down_proj).  Embedding, norms, the vision tower and lm_head are replicated.
# ------------------------------------------------------------------------------------------------------------------
def shard_llm_weights(weights: dict, geo, rank: int, world: int):
    """(weights of rank `rank`, its local Geometry).  `weights`: HF parameter names -> full tensors (not modified)."""
    from dataclasses import replace
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    if geo.num_kv_heads % world or geo.num_heads % world or geo.mlp % world:
        raise ValueError(f"kv heads {geo.num_kv_heads}, q heads {geo.num_heads} and mlp width {geo.mlp} must be divisible by the TP degree {world}")
    d = geo.head_dim
    q0, q1 = rank * geo.num_heads // world * d, (rank + 1) * geo.num_heads // world * d
    k0, k1 = rank * geo.num_kv_heads // world * d, (rank + 1) * geo.num_kv_heads // world * d
    m0, m1 = rank * geo.mlp // world, (rank + 1) * geo.mlp // world
    out = {}
    for name, t in weights.items():
        if ".self_attn.q_proj." in name:
            t = t[q0:q1]
        elif ".self_attn.k_proj." in name or ".self_attn.v_proj." in name:
            t = t[k0:k1]
        elif name.endswith(".self_attn.o_proj.weight"):
            t = t[:, q0:q1]
        elif "language_model" in name and (".mlp.gate_proj." in name or ".mlp.up_proj." in name):
            t = t[m0:m1]
        elif "language_model" in name and name.endswith(".mlp.down_proj.weight"):
            t = t[:, m0:m1]
        out[name] = t.contiguous()
    local = replace(geo, num_heads=geo.num_heads // world, num_kv_heads=geo.num_kv_heads // world, mlp=geo.mlp // world)
    return out, local


class TpContext:
    """Symmetric-memory state of one tensor-parallel group on this rank (`svlm_tp_ctx_t`): one buffer of the same size on every
    rank, every peer's copy mapped into this process over NVLink (`torch.distributed._symmetric_memory`: CUDA VMM + fabric / fd
    handles; plumbing only -- the exchange itself is `svlm_tp_gemv_allreduce` / `svlm_tp_allreduce_rows`, no NCCL call on the
    data path).  Collective: every rank of `group` must construct it at the same time."""

    def __init__(self, group, hidden: int, max_rows: int = 1024, device=None):
        import ctypes

        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import _abi
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= _abi.MAX_TP:
            raise ValueError(f"tensor-parallel degree must be 2..{_abi.MAX_TP}, got {self.world}")
        dev = torch.device(device if device is not None else torch.cuda.current_device())
        self.hidden, self.max_rows = hidden, max_rows
        self.region = -(-max(self.world * hidden * 8, max_rows * hidden * 2) // 1024) * 1024
        nbytes = _abi.TP_DECODE_OFFSET + 4 * self.region
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        self.handle = symm.rendezvous(self.buf, group)
        self.buf.zero_()
        self.counters = torch.zeros(2, dtype=torch.int32, device=dev)        # seq, done
        torch.cuda.synchronize(dev)
        dist.barrier(group)                                                  # every rank's flags are zero before anyone signals
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected peer table")
        self.c = _abi.TpCtx()
        self.c.world, self.c.rank = self.world, self.rank
        for r, p_ in enumerate(ptrs):
            self.c.peer_base[r] = p_
        self.c.region_bytes = self.region
        self.c.seq = self.counters[0:1].data_ptr()
        self.c.done = self.counters[1:2].data_ptr()
        self.ref = ctypes.byref(self.c)
        self.all_reduces = 0

    def rows_view(self, which: int, rows: int, cols: int):
        """bf16 [rows, cols] view of region `which` of this rank's row-partial area: the GEMM writes its partial products here."""
        from . import _abi
        if rows * cols * 2 > self.region:
            raise RuntimeError(f"{rows} x {cols} partial products do not fit a {self.region}-byte region (max_rows = {self.max_rows})")
        off = _abi.TP_DECODE_OFFSET + (2 + which) * self.region
        return self.buf[off:off + rows * cols * 2].view(torch.bfloat16).view(rows, cols)

    def nvlink_bytes_per_decode_allreduce(self):
        """Bytes this rank pushes over NVLink in one decode all-reduce: (world - 1) peers x hidden (fp32 value, epoch) words."""
        return (self.world - 1) * self.hidden * 8
