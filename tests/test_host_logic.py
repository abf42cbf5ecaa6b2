"""CPU: the product's host-side logic (range finder, eviction planner, run encoding) against the
fixtures produced by the reference's own functions, independently of the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from streaming_vlm_b200.inference import inference as host
from streaming_vlm_b200.utils.get_qwen_range import get_qwen_range
from tests.stream_sim import replay, rle, sha


class TagCache:
    """Stands in for the arena-backed StreamingCache: records the gather `compact` is asked for."""

    def __init__(self, n):
        self.n = n
        self.gather = list(range(n))

    def get_seq_length(self):
        return self.n

    def compact(self, runs, new_ids):
        d = 0
        g = []
        for s, dst, l in runs:
            assert dst == d
            g += list(range(s, s + l))
            d += l
        self.gather = g
        self.n = d


@pytest.fixture(scope="module")
def streams(golden_dir):
    return json.load(open(os.path.join(golden_dir, "eviction_streams.json")))


def test_get_qwen_range_matches_reference(golden_dir, streams):
    probes = json.load(open(os.path.join(golden_dir, "qwen_range_probes.json")))
    by_round = {(si, r["i"]): r["after_ids"] for si, item in enumerate(streams) for r in item["rounds"] if "after_ids" in r}
    for p in probes:
        ids = p["ids"] if "adv" in p else by_round[(p["stream"], p["round"])]
        try:
            got = list(get_qwen_range(torch.tensor([ids]), p["label"], p["index"], p["contain_lf"]))
        except IndexError:
            got = None
        assert got == p["range"], p


@pytest.mark.parametrize("si", range(5))
def test_process_past_kv_matches_reference(streams, si):
    item = streams[si]
    cfg = item["config"]
    n = 0
    max_runs = 0
    for rec, before, after, gather in replay(host.process_past_kv, TagCache, lambda c: c.gather, cfg, item["rounds"]):
        assert sha(after) == rec["after_ids_sha"], (cfg["name"], rec["i"])
        assert rle(gather) == rec["cache_runs"], (cfg["name"], rec["i"])
        max_runs = max(max_runs, len(rec["cache_runs"]))
        n += 1
    assert n == cfg["rounds"] - 1
    assert max_runs <= 8          # SVLM_MAX_RUNS is 32


def test_prune_interval():
    ids = torch.arange(20).view(1, -1)
    c = TagCache(18)
    new_ids, c = host.prune_id_and_kv_cache(ids, c, 3, 6)
    assert new_ids.flatten().tolist() == [0, 1, 2] + list(range(7, 20))
    assert c.gather == [0, 1, 2] + list(range(7, 18))


# ------------------------------------------------------------------------------------------------ property test of the range finder
def test_find_segments_equals_oracle_on_random_token_soups():
    """The vectorised range finder of the product against the oracle's sequential restatement of `get_qwen_range` on random
    sequences over the structural tokens (nested, unterminated and overlapping patterns included), every label, both
    `contain_lf` settings."""
    from hypothesis import given, settings, strategies as st
    from oracle import tokens as o_tok
    from streaming_vlm_b200.utils.get_qwen_range import TOKEN_IDS, find_segments
    alphabet = [TOKEN_IDS["<|im_start|>"], TOKEN_IDS["<|im_end|>"], TOKEN_IDS["user"], TOKEN_IDS["assistant"], TOKEN_IDS["Time"],
                TOKEN_IDS["\n"], TOKEN_IDS["<|vision_start|>"], TOKEN_IDS["<|vision_end|>"], TOKEN_IDS["<|video_pad|>"],
                *TOKEN_IDS["previous text"], 1000, 1001]

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.sampled_from(alphabet), min_size=0, max_size=60))
    def check(ids):
        arr = np.asarray(ids, dtype=np.int64)
        for label in ("user", "previous text", "user_text", "assistant", "vision"):
            for lf in (True, False):
                want = [tuple(x) for x in o_tok.all_ranges(ids, label, lf)]
                got = find_segments(arr, label, lf)
                assert got == want, (label, lf, ids)

    check()
