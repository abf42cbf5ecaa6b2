"""CPU, world_size 2 over gloo: the N > 1 host logic -- stream partitioning and the throughput aggregation
bench.py uses (units summed over ranks, time = max over ranks); every rank plans its own stream's eviction
independently and the plans must equal the single-process plans (no cross-rank state)."""
import json
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streaming_vlm_b200 import parallel, synthetic


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.assign_streams(5, rank, world)
    states = {s: synthetic.fast_forward(40, seed=s, n_vision_tokens=16, grid=(1, 8, 8)) for s in mine}
    units, secs = parallel.aggregate(len(mine) * 10.0, 1.0 + rank)
    json.dump({"mine": mine, "units": units, "secs": secs, "lens": {str(s): st["cache_len"] for s, st in states.items()}},
              open(os.path.join(out_dir, f"r{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_sharding(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [json.load(open(tmp_path / f"r{k}.json")) for k in range(2)]
    assert r[0]["mine"] == [0, 2, 4] and r[1]["mine"] == [1, 3]
    assert sorted(r[0]["mine"] + r[1]["mine"]) == list(range(5))
    for k in range(2):
        assert r[k]["units"] == 50.0 and r[k]["secs"] == 2.0          # sum of units, max of time
    for s in range(5):
        want = synthetic.fast_forward(40, seed=s, n_vision_tokens=16, grid=(1, 8, 8))["cache_len"]
        assert r[s % 2]["lens"][str(s)] == want


# ------------------------------------------------------------------------------------------------ head-sharded tensor parallelism
def _tp_worker(rank, world, port, out_dir):
    """Every rank runs the ORACLE decoder layer on its weight shard (fp32, on CPU); the row-parallel partial products are
    summed with a gloo all-reduce at the two points the CUDA engine reduces at.  The result must equal the unsharded layer."""
    import torch.nn.functional as F
    from oracle.attention import attention
    from oracle.positions import apply_mrope, default_inv_freq, rotary_cos_sin
    from oracle.vit import rmsnorm
    from streaming_vlm_b200.config import TINY
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    geo = TINY
    w = {k: v.float() for k, v in synthetic.random_weights(geo, seed=3).items()}
    mine, local = parallel.shard_llm_weights(w, geo, rank, world)
    assert local.num_heads == geo.num_heads // world and local.num_kv_heads == geo.num_kv_heads // world and local.mlp == geo.mlp // world
    t = 9
    x = torch.randn(t, geo.hidden, generator=torch.Generator().manual_seed(5))
    pos = torch.arange(t, dtype=torch.float32).view(1, 1, t).expand(3, 1, t)
    cos, sin = rotary_cos_sin(pos, default_inv_freq(geo.head_dim, geo.rope_theta), torch.float32)

    def layer(wd, g, reduce):
        p = "model.language_model.layers.0."
        h = rmsnorm(x, wd[p + "input_layernorm.weight"], g.rms_eps)
        q = F.linear(h, wd[p + "self_attn.q_proj.weight"], wd[p + "self_attn.q_proj.bias"]).view(1, t, g.num_heads, g.head_dim).transpose(1, 2)
        k = F.linear(h, wd[p + "self_attn.k_proj.weight"], wd[p + "self_attn.k_proj.bias"]).view(1, t, g.num_kv_heads, g.head_dim).transpose(1, 2)
        v = F.linear(h, wd[p + "self_attn.v_proj.weight"], wd[p + "self_attn.v_proj.bias"]).view(1, t, g.num_kv_heads, g.head_dim).transpose(1, 2)
        q, k = apply_mrope(q, k, cos, sin, g.mrope_section)
        o = attention(q[0], k[0], v[0], causal=True)
        y = x + reduce(F.linear(o, wd[p + "self_attn.o_proj.weight"]))
        h = rmsnorm(y, wd[p + "post_attention_layernorm.weight"], g.rms_eps)
        act = F.silu(F.linear(h, wd[p + "mlp.gate_proj.weight"])) * F.linear(h, wd[p + "mlp.up_proj.weight"])
        return y + reduce(F.linear(act, wd[p + "mlp.down_proj.weight"]))

    def all_reduce(z):
        z = z.contiguous()
        dist.all_reduce(z)
        return z

    got = layer(mine, local, all_reduce)
    want = layer(w, geo, lambda z: z)
    torch.save({"err": float((got - want).abs().max()), "scale": float(want.abs().max())}, os.path.join(out_dir, f"tp{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tensor_parallel_sharding(tmp_path):
    port = 31500 + os.getpid() % 2000
    mp.spawn(_tp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for k in range(2):
        r = torch.load(tmp_path / f"tp{k}.pt")
        assert r["err"] < 1e-4 * max(r["scale"], 1.0), r
