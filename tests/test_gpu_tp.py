"""GPU (needs >= 2 devices on the box, skipped otherwise): head-sharded tensor parallelism over NVLink symmetric memory.  Two ranks
(torchrun) run the tiny stream with the language model sharded; every rank is teacher-forced with the tokens of an unsharded engine
and must reproduce its ids / cache lengths, the ranks must agree with each other bit for bit, and the steps run as CUDA-graph
replays (the all-reduce is `svlm_tp_gemv_allreduce` / `svlm_tp_allreduce_rows`, not NCCL).  See benchmarks/tp_parity_and_bench.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="tensor parallelism needs two GPUs on the box")
def test_two_rank_tensor_parallel_stream_matches_unsharded():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "benchmarks", "tp_parity_and_bench.py"), "parity"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 2 and all(l["cuda_graphs"] and l["tp"] == 2 for l in lines)
    assert all(l["argmax_mismatches_vs_unsharded"] <= 12 for l in lines)
