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
