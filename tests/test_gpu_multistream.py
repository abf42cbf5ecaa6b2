"""GPU parity of the multi-stream engine (several streams per GPU, BASELINE config 4 at one GPU): B tiny-model streams
with different frames and prompt lengths advance through ONE batched decode step per token; each stream is checked
against its own CPU oracle session exactly like the single-stream test (`tests/test_gpu_session.py`): pruned ids, gather
runs and cache lengths bit-exact, greedy ids identical up to oracle near-ties (teacher-forced)."""
import pytest
import torch

from oracle.model import Session
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import TINY
from tests.parity_log import MAX_NEAR_TIE_FRACTION, near_tie, record
from tests.stream_sim import rle

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.08


def _run(n_streams, n_chunks, height, width, window, text_round, sink, sliding, use_graph=True, batch_prefill=True):
    from streaming_vlm_b200.multi_stream import MultiStreamEngine
    geo = TINY
    w_cpu = synthetic.random_weights(geo, seed=0)
    suppress = synthetic.suppress_mask(geo.vocab)
    oracles = [Session(w_cpu, geo.as_dict(), window, text_round, sink, sliding, suppress=suppress) for _ in range(n_streams)]
    multi = MultiStreamEngine({k: v.clone() for k, v in w_cpu.items()}, geo, n_streams, capacity=4096, window_size=window,
                              text_round=text_round, text_sink=sink, text_sliding_window=sliding, suppress=suppress, use_graph=use_graph)
    multi.batch_prefill = batch_prefill
    n_vis = (height // 28) * (width // 28)
    mismatches = total = 0
    prev_len = [0] * n_streams
    for i in range(n_chunks):
        prompts, pixels, ros = [], [], []
        for b in range(n_streams):
            px, grid = synthetic.patchify(synthetic.frames(1000 * b + i, height, width), geo)
            prompt = synthetic.round_prompt(i, n_vis, 4 + 3 * b)          # different previous-text lengths per stream
            ros.append(oracles[b].step(prompt, px.to(torch.bfloat16), grid))
            prompts.append(prompt)
            pixels.append(px.to("cuda"))
        rgs = multi.step(prompts, pixels, grid, forced=[r["new_tokens"] for r in ros])
        for b, (rg, ro) in enumerate(zip(rgs, ros)):
            assert rg["ids"] == ro["ids"], f"chunk {i} stream {b}: ids"
            assert rg["input_len"] == ro["input_len"] and rg["cache_len"] == ro["cache_len"], f"chunk {i} stream {b}"
            if ro["gather"] is not None:
                want = [x for x in ro["gather"] if x < prev_len[b]]
                got = [s_ + k for s_, _, l in rg["runs"] for k in range(l)]
                assert got == want, f"chunk {i} stream {b}: gather {rle(got)} vs {rle(want)}"
            prev_len[b] = ro["cache_len"]
            for j, (tg, to) in enumerate(zip(multi.last_argmax[b], ro["new_tokens"])):
                total += 1
                if tg != to:
                    mismatches += 1
                    lo = ro["step_logits"][j]
                    assert near_tie(lo, tg, to), f"chunk {i} stream {b} token {j}: engine {tg} oracle {to} margin {lo[to] - lo[tg]} top {lo[to]}"
    return mismatches, total, multi


@pytest.mark.parametrize("batch_prefill", [True, False])
def test_three_streams_small_windows(batch_prefill):
    """batch_prefill=True: ViT and chunk prefill of all streams in one pass; False: per stream (the single-stream path)."""
    mism, total, multi = _run(3, 12, 112, 112, 4, 4, 16, 24, batch_prefill=batch_prefill)
    assert all(e.cache.compactions >= 6 for e in multi.engines)
    assert len({tuple(e.prev_ids.tolist()) for e in multi.engines}) == 3          # the streams really differ
    record("multistream", batch_prefill=batch_prefill, greedy_mismatches=mism, greedy_tokens=total)
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_ten_streams_one_batched_step():
    """More than 8 streams per launch (SVLM_MAX_STREAMS = 16): 10 streams through the batched ViT / prefill / decode, each against its
    own oracle session."""
    mism, total, multi = _run(10, 5, 112, 112, 4, 4, 16, 24)
    assert len({tuple(e.prev_ids.tolist()) for e in multi.engines}) == 10
    record("multistream_10", greedy_mismatches=mism, greedy_tokens=total)
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_two_streams_eager_equals_graph():
    """The CUDA-graph replay of the batched step and its eager execution produce the same tokens."""
    _, _, a = _run(2, 6, 112, 112, 4, 4, 16, 24, use_graph=True)
    _, _, b = _run(2, 6, 112, 112, 4, 4, 16, 24, use_graph=False)
    for ea, eb in zip(a.engines, b.engines):
        assert ea.prev_ids.tolist() == eb.prev_ids.tolist()
        assert ea.cache.seq_len == eb.cache.seq_len


def test_multistream_lookahead_loop_equals_plain_loop():
    """The batched decode loop with one step of look-ahead (next step enqueued before the host has read the current tokens; streams
    whose turn ends are switched off ahead of it or rolled back after it) == the plain read-then-step loop: same ids per stream, same
    cache lengths, device-side counters and arenas bit for bit.  Four-id vocabulary incl. <|im_end|>: turns end early all the time,
    at different steps in different streams."""
    from streaming_vlm_b200.engine import IM_END
    from streaming_vlm_b200.multi_stream import MultiStreamEngine
    geo = TINY
    w_cpu = synthetic.random_weights(geo, seed=0)
    suppress = torch.ones(geo.vocab, dtype=torch.bool)
    suppress[torch.tensor([IM_END, 1000, 1001, 1002])] = False
    multis = []
    for look in (True, False):
        m = MultiStreamEngine({k: v.clone() for k, v in w_cpu.items()}, geo, 3, capacity=4096, window_size=4, text_round=4, text_sink=16,
                              text_sliding_window=24, suppress=suppress, prefill_bucket=1)
        m.lookahead = look
        multis.append(m)
    early = 0
    for i in range(12):
        prompts, pixels = [], []
        for b in range(3):
            px, grid = synthetic.patchify(synthetic.frames(1000 * b + i, 112, 112), geo)
            prompts.append(synthetic.round_prompt(i, 16, 4 + 3 * b))
            pixels.append(px.to("cuda"))
        ra, rb = (m.step(prompts, pixels, grid) for m in multis)
        for b in range(3):
            assert ra[b]["ids"] == rb[b]["ids"], f"chunk {i} stream {b}"
            assert ra[b]["cache_len"] == rb[b]["cache_len"] and ra[b]["runs"] == rb[b]["runs"], f"chunk {i} stream {b}"
            early += len(ra[b]["new_tokens"]) < multis[0].engines[b].max_new_tokens
        for name in ("kv_len", "next_pos"):
            assert torch.equal(getattr(multis[0], name), getattr(multis[1], name)), f"chunk {i}: {name}"
    assert early >= 5, "the test streams never ended a turn early"
    for b in range(3):
        ca, cb = multis[0].engines[b].cache, multis[1].engines[b].cache
        n = ca.seq_len
        assert torch.equal(ca.arena[ca.cur, :, :, :, :n].view(torch.int16), cb.arena[cb.cur, :, :, :, :n].view(torch.int16)), f"stream {b}: arenas"
