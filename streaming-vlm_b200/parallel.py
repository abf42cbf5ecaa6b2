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
