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


def test_native_find_segment_equals_vectorised_scan_and_oracle():
    """`svlm_find_segment` (host function of the C ABI, what the planner calls) against the numpy scan and the oracle's sequential
    restatement: every label, both `contain_lf` settings, every valid index from both ends, IndexError beyond them -- on random token
    soups (nested, unterminated, overlapping patterns) and on a long steady-state stream."""
    from hypothesis import given, settings, strategies as st
    from oracle import tokens as o_tok
    from streaming_vlm_b200 import synthetic
    from streaming_vlm_b200.utils.get_qwen_range import TOKEN_IDS, find_segment, find_segments
    alphabet = [TOKEN_IDS["<|im_start|>"], TOKEN_IDS["<|im_end|>"], TOKEN_IDS["user"], TOKEN_IDS["assistant"], TOKEN_IDS["Time"],
                TOKEN_IDS["\n"], TOKEN_IDS["<|vision_start|>"], TOKEN_IDS["<|vision_end|>"], TOKEN_IDS["<|video_pad|>"],
                *TOKEN_IDS["previous text"], 1000, 1001]

    def compare(ids_list):
        arr = np.asarray(ids_list, dtype=np.int64)
        for label in ("user", "previous text", "user_text", "assistant", "vision"):
            for lf in (True, False):
                want = [tuple(x) for x in o_tok.all_ranges(list(ids_list), label, lf)]
                assert find_segments(arr, label, lf) == want
                n = len(want)
                for index in list(range(n)) + list(range(-n, 0)):
                    assert find_segment(arr, label, index, lf) == want[index], (label, lf, index)
                for index in (n, -n - 1):
                    with pytest.raises(IndexError):
                        find_segment(arr, label, index, lf)

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.sampled_from(alphabet), min_size=0, max_size=60))
    def check(ids):
        compare(ids)

    check()
    st_ = synthetic.fast_forward(40, new_tokens=20, seed=3)
    ids = np.asarray(st_["prev_ids"], dtype=np.int64)
    for label in ("user", "previous text", "user_text", "assistant", "vision"):
        segs = find_segments(ids, label, True)
        assert len(segs) >= 1
        for index in (0, -1, len(segs) // 2, -len(segs)):
            assert find_segment(ids, label, index, True) == segs[index]


def test_planner_equals_oracle_on_random_policies():
    """Product planner (`plan_past_kv`, numpy) against the oracle's `plan_eviction` (pinned on five policies) over 40 RANDOM
    policies -- window, text rounds, sink and sliding-window sizes, turn lengths, early <|im_end|> -- 24 rounds each: pruned ids
    and gather maps must agree everywhere the oracle accepts the configuration, and both must refuse where it does not."""
    import copy
    from oracle import eviction as o_ev
    from streaming_vlm_b200 import synthetic
    from streaming_vlm_b200.inference.inference import plan_past_kv
    rng = np.random.RandomState(2024)
    compared = refused = 0
    for _ in range(40):
        window, text_round = int(rng.randint(1, 7)), int(rng.randint(1, 7))
        sink = None if rng.rand() < 0.2 else int(rng.randint(2, 40))
        slide = None if rng.rand() < 0.2 else int(rng.randint(2, 40))
        n_vis, n_prev, seed = int(rng.choice([4, 16])), int(rng.randint(2, 30)), int(rng.randint(0, 1000))
        hist_o, clips_o, px_o = [], [], []
        hist_p, clips_p, px_p = [], [], []
        prev = None
        for i in range(24):
            if prev is not None:
                cache_len = len(prev) - 1
                try:
                    want_ids, gather = o_ev.plan_eviction(list(prev), i, text_round, window, hist_o, 3, 2, clips_o, px_o, sink, slide)
                    ok = True
                except (AssertionError, IndexError):
                    ok = False
                if not ok:
                    with pytest.raises((AssertionError, IndexError)):
                        plan_past_kv(np.asarray(prev), i, text_round, window, hist_p, 3, 2, clips_p, px_p, sink, slide)
                    refused += 1
                    break
                plan = plan_past_kv(np.asarray(prev), i, text_round, window, hist_p, 3, 2, clips_p, px_p, sink, slide)
                assert plan.ids.tolist() == want_ids, (window, text_round, sink, slide, i)
                got = [s_ + k for s_, _, l in plan.runs(cache_len) for k in range(l)]
                assert got == [g for g in gather if g < cache_len], (window, text_round, sink, slide, i)
                assert hist_p == hist_o and len(clips_p) == len(clips_o)
                prev = want_ids
                compared += 1
            for clips, px, hist in ((clips_o, px_o, hist_o), (clips_p, px_p, hist_p)):
                clips.append(None)
                px.append(None)
                if i == 0:
                    hist[:] = [{"role": "previous text", "content": "p" * 10},
                               {"role": "user", "content": [{"type": "text", "text": "T"}, {"type": "video"}, {"type": "text", "text": "q"}]}]
                else:
                    hist.append({"role": "user", "content": [{"type": "text", "text": "T"}, {"type": "video"}]})
            prompt = synthetic.round_prompt(i, n_vis, n_prev, 8, seed)
            ids = prompt if prev is None else list(prev) + (prompt if prev[-1] != 198 else prompt[1:])
            gen = [int(x) for x in rng.randint(1000, 100000, size=int(rng.randint(1, 13)))]
            gen = [g if g not in (872, 77091, 1462, 198, 19702, 1467) else 1000 for g in gen]
            if rng.rand() < 0.3:
                gen[-1] = 151645
            seq = ids + gen
            if seq[-1] != 151645:
                seq = seq + [151645]
            prev = seq
            for hist in (hist_o, hist_p):
                hist.append({"role": "assistant", "content": "x" * 8 + " ..."})
    assert compared > 500, (compared, refused)
