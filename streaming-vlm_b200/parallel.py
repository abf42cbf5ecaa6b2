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
