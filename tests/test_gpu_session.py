"""GPU parity, BASELINE config 1: tiny random-init Qwen2.5-VL-architecture model streaming 64 synthetic
448x448 frames (32 one-second chunks: 16 without eviction, 16 with) through sink / text-window /
vision-window eviction, CUDA engine vs the CPU oracle in lock step.

Per chunk:  gather runs, pruned ids and position ids bit-exact; K_rot bit-exact w.r.t. the reference
mRoPE of the arena's own K_raw; per-layer attention outputs and last-position logits within a bf16
tolerance; greedy token ids identical, where a differing argmax is tolerated only if the oracle's own
margin between the two candidates is below the logit tolerance (the stream is then teacher-forced with
the oracle's token, the reference's --gt_json mechanism, so that one near-tie cannot fork the stream).
"""
import numpy as np
import pytest
import torch

from oracle import positions as o_pos
from oracle.model import Session
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import TINY
from tests.parity_log import MAX_NEAR_TIE_FRACTION, near_tie, record
from tests.stream_sim import rle

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.08        # |logit| ~ 1; bf16 hidden state (8 mantissa bits) through a 512-wide lm_head
ATTN_TOL = dict(rtol=3e-2, atol=2e-2)


def _run(n_chunks, height, width, window, text_round, sink, sliding, n_prev, check_every=1, pos_mode="shrink"):
    from streaming_vlm_b200.engine import StreamingEngine
    geo = TINY
    w_cpu = synthetic.random_weights(geo, seed=0)
    suppress = synthetic.suppress_mask(geo.vocab)
    oracle = Session(w_cpu, geo.as_dict(), window, text_round, sink, sliding, suppress=suppress, pos_mode=pos_mode)
    engine = StreamingEngine({k: v.clone() for k, v in w_cpu.items()}, geo, capacity=8192, window_size=window,
                             text_round=text_round, text_sink=sink, text_sliding_window=sliding, suppress=suppress, pos_mode=pos_mode)
    n_vis = (height // 28) * (width // 28)
    mismatches, total, prev_cache_len = 0, 0, 0
    ve_worst = 0.0
    for i in range(n_chunks):
        px, grid = synthetic.patchify(synthetic.frames(i, height, width), geo)
        prompt = synthetic.round_prompt(i, n_vis, n_prev)
        taps_o, taps_g = {}, {}
        deep = (i % check_every == 0) or i == n_chunks - 1
        ro = oracle.step(prompt, px.to(torch.bfloat16), grid, taps=taps_o if deep else None)
        rg = engine.step(prompt, px.to("cuda"), grid, taps=taps_g if deep else None, forced=ro["new_tokens"])
        # ---- integer state: bit-exact
        assert rg["ids"] == ro["ids"], f"chunk {i}: ids"
        assert rg["input_len"] == ro["input_len"] and rg["cache_len"] == ro["cache_len"], f"chunk {i}"
        if ro["gather"] is not None:
            want = [x for x in ro["gather"] if x < prev_cache_len]
            got = [s_ + k for s_, _, l in rg["runs"] for k in range(l)]
            assert got == want, f"chunk {i}: gather {rle(got)} vs {rle(want)}"
        prev_cache_len = ro["cache_len"]
        if deep:
            n = rg["input_len"]
            pos_ref = taps_o["position_ids"][:, 0, :].numpy()
            pos_got = engine.cache.pos[:, :n].cpu().numpy().astype(np.float32)
            if pos_mode == "append":
                # the oracle (like the reference) only ever computes the positions of the NEW tokens; retained tokens keep theirs
                t_new = pos_ref.shape[1]
                assert np.array_equal(pos_got[:, n - t_new:n], pos_ref), f"chunk {i}: positions of the new tokens"
                idx = engine.cache.pos[:, :n].cpu().long()
            else:
                # engine positions cover ids of the prefill; decode tokens continue the text run
                assert np.array_equal(pos_got, pos_ref[:, :n]), f"chunk {i}: positions"
                idx = taps_o["position_ids"][:, 0, :n].long()
            # K_rot == reference mRoPE applied to the arena's own K_raw at those positions (bit-exact); cos/sin gathered from
            # the engine's own host-built table (same arithmetic as HF's rotary module, see
            # tests/test_oracle_golden.py::test_rotary_table_equals_per_position_embedding)
            k_raw, k_rot, _ = engine.cache.planes(0)
            tc, ts = engine.cache.cos_table.cpu()[idx], engine.cache.sin_table.cpu()[idx]
            cos, sin = torch.cat((tc, tc), dim=-1)[:, None], torch.cat((ts, ts), dim=-1)[:, None]
            kr = k_raw[:, :n].cpu()
            _, ref_rot = o_pos.apply_mrope(kr[None, :1, :1], kr[None], cos, sin, geo.mrope_section)
            assert torch.equal(k_rot[:, :n].cpu().view(torch.int16), ref_rot[0].view(torch.int16)), f"chunk {i}: K_rot"
            # floating point rows
            ve = (taps_g["video_embeds"].float().cpu() - taps_o["video_embeds"].float()).abs().max().item()
            ve_worst = max(ve_worst, ve / taps_o["video_embeds"].float().abs().max().item())
            torch.testing.assert_close(taps_g["video_embeds"].float().cpu(), taps_o["video_embeds"].float(), rtol=3e-2, atol=3e-2)
            for li in range(geo.num_layers):
                torch.testing.assert_close(taps_g["attn_out"][li].float().cpu(), taps_o["attn_out"][li].float(), **ATTN_TOL)
            lg, lo = taps_g["logits"].cpu(), taps_o["logits"]
            keep = ~suppress
            assert (lg[keep] - lo[keep]).abs().max() < LOGIT_TOL, f"chunk {i}: logits {(lg[keep] - lo[keep]).abs().max()}"
        # ---- greedy ids
        for j, (tg, to) in enumerate(zip(engine.last_argmax, ro["new_tokens"])):
            total += 1
            if tg != to:
                mismatches += 1
                lo = ro["step_logits"][j]
                assert near_tie(lo, tg, to), f"chunk {i} token {j}: engine {tg} oracle {to} margin {lo[to] - lo[tg]} top {lo[to]}"
    record("tiny_stream", chunks=n_chunks, frame=f"{height}x{width}", pos_mode=pos_mode, greedy_mismatches=mismatches, greedy_tokens=total,
           video_embeds_max_rel=ve_worst)
    return mismatches, total, engine, oracle


def test_tiny_stream_small_windows():
    """Fast: 112x112 frames (16 vision tokens / chunk), window 4, text_round 4, sink 16, window 24 -> every
    eviction branch (move, user block drop, previous-text trim) is exercised within 20 chunks."""
    mism, total, engine, _ = _run(20, 112, 112, 4, 4, 16, 24, 6)
    assert engine.cache.compactions >= 14
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)       # near-ties only: each one was checked against the oracle's margin


def test_tiny_stream_config1():
    """BASELINE.json configs[0]: 64 frames of 448x448 = 32 chunks, defaults 16/16/512/512."""
    mism, total, engine, oracle = _run(32, 448, 448, 16, 16, 512, 512, 24, check_every=4)
    assert engine.cache.compactions >= 15
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_tiny_stream_append_position_mode():
    """`pos_mode="append"` (StreamingArgs, model_forward.py:112-153): positions keep growing while the cache is pruned; retained
    keys keep the rotation they were appended with.  Same checks as the shrink-mode stream, every eviction branch included."""
    mism, total, engine, oracle = _run(20, 112, 112, 4, 4, 16, 24, 6, pos_mode="append")
    assert engine.cache.compactions >= 14
    assert int(engine.cache.next_pos) == oracle.pos_state["last_cache_position"] + 1
    assert int(engine.cache.next_pos) > 2 * engine.cache.seq_len          # positions really outgrew the cache
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_lookahead_decode_equals_plain_loop():
    """The greedy loop with one step of look-ahead (next decode step enqueued before the host has seen the current token,
    rolled back when that token is <|im_end|>) produces the same stream as the plain read-then-step loop.  The vocabulary is
    narrowed to four ids incl. <|im_end|> so that turns end early all the time; both position modes."""
    from streaming_vlm_b200.engine import IM_END, StreamingEngine
    geo = TINY
    w_cpu = synthetic.random_weights(geo, seed=0)
    suppress = torch.ones(geo.vocab, dtype=torch.bool)
    suppress[torch.tensor([IM_END, 1000, 1001, 1002])] = False
    for pos_mode in ("shrink", "append"):
        engines = [StreamingEngine({k: v.clone() for k, v in w_cpu.items()}, geo, capacity=4096, window_size=4, text_round=4, text_sink=16,
                                   text_sliding_window=24, suppress=suppress, use_graph=ug, pos_mode=pos_mode, prefill_bucket=1)
                   for ug in (True, False)]         # prefill_bucket=1: same GEMM shapes on both sides, so the arenas can be compared bit for bit
        early = 0
        for i in range(14):
            px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
            prompt = synthetic.round_prompt(i, 16, 6)
            ra, rb = (e.step(prompt, px.to("cuda"), grid) for e in engines)
            assert ra["ids"] == rb["ids"], f"{pos_mode} chunk {i}"
            assert ra["cache_len"] == rb["cache_len"] and ra["runs"] == rb["runs"], f"{pos_mode} chunk {i}"
            assert int(engines[0].cache.next_pos) == int(engines[1].cache.next_pos), f"{pos_mode} chunk {i}: next position"
            early += len(ra["new_tokens"]) < engines[0].max_new_tokens
        assert early >= 3, "the test stream never ended a turn early"
        n = engines[0].cache.seq_len
        a, b = (e.cache.arena[e.cache.cur, :, :, :, :n] for e in engines)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), f"{pos_mode}: arenas differ"


def test_padded_prefill_graphs_and_variable_turns_stay_bounded():
    """Chunk-prefill graphs are keyed on a token-count BUCKET (padded tail rows, device-side true length) and both graph caches are
    LRU-bounded: a stream whose turns end after 1..20 tokens and whose prompts vary in length (1..9 time tokens) must (a) give
    the same stream as exact-length prefills, (b) keep the number of captured graphs and the allocated HBM flat."""
    import gc
    from streaming_vlm_b200.engine import IM_END, StreamingEngine
    gc.collect()                                # engines of earlier tests must not be freed in the middle of the HBM-flatness window
    torch.cuda.synchronize()
    geo = TINY
    w_cpu = synthetic.random_weights(geo, seed=0)
    suppress = torch.ones(geo.vocab, dtype=torch.bool)
    suppress[torch.tensor([IM_END] + list(range(1000, 1012)))] = False          # 13 ids: turns end early at random points
    mk = lambda bucket, limit: StreamingEngine({k: v.clone() for k, v in w_cpu.items()}, geo, capacity=4096, window_size=4, text_round=4,
                                               text_sink=16, text_sliding_window=24, suppress=suppress, prefill_bucket=bucket, max_graphs=limit)
    padded, exact = mk(8, 6), mk(1, 64)
    lens, mem, worst = set(), [], 0.0
    for i in range(60):
        px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
        prompt = synthetic.round_prompt(i, 16, 6, time_tokens=1 + (i * 5) % 9)
        ra = padded.step(prompt, px.to("cuda"), grid)
        rb = exact.step(prompt, px.to("cuda"), grid, forced=ra["new_tokens"])
        assert ra["ids"] == rb["ids"] and ra["cache_len"] == rb["cache_len"] and ra["runs"] == rb["runs"], f"chunk {i}"
        lens.add(len(ra["new_tokens"]))
        n = padded.cache.seq_len
        a, b = (e.cache.arena[e.cache.cur, :, :, :, :n].float() for e in (padded, exact))
        worst = max(worst, ((a - b).abs().max() / b.abs().max()).item())
        # the exact engine's own selections: identical except at near-ties of the two GEMM shapes
        mem.append(torch.cuda.memory_allocated())
    record("padded_prefill_variable_turns", turn_lengths=sorted(lens), prefill_graphs=len(padded._pf_graphs), decode_graphs=len(padded._graphs),
           prefill_graph_evictions=padded._pf_graphs.evictions, exact_prefill_graphs=len(exact._pf_graphs), arena_max_rel_diff=worst)
    assert len(lens) >= 5, lens                                            # the turns really varied
    assert worst < 2e-2, worst                                             # same K/V rows up to the GEMM-shape rounding
    assert len(padded._pf_graphs) <= 6 and len(padded._graphs) <= 6
    assert len(exact._pf_graphs) > len(padded._pf_graphs) or padded._pf_graphs.evictions == 0
    assert max(mem[30:]) - min(mem[30:]) < (64 << 20), "allocated HBM keeps moving with bounded graph caches"


@pytest.mark.parametrize("seed", [1, 3, 4])
def test_unforced_stream_identical_greedy_ids(seed):
    """Fixed seeds, fixed horizon, NO teacher forcing: the engine's free-running greedy stream equals the oracle's token for token
    over 10 chunks (70 generated tokens, every eviction branch).  The seeds are those of `benchmarks/scan_unforced_seeds.py` whose
    streams contain no bf16 near-tie (other seeds fork at one; the teacher-forced tests above bound and record those)."""
    from streaming_vlm_b200.engine import StreamingEngine
    geo = TINY
    w = synthetic.random_weights(geo, seed=seed)
    sup = synthetic.suppress_mask(geo.vocab)
    oracle = Session(w, geo.as_dict(), 4, 4, 16, 24, suppress=sup, max_new_tokens=6)
    engine = StreamingEngine({k: v.clone() for k, v in w.items()}, geo, capacity=2048, window_size=4, text_round=4, text_sink=16,
                             text_sliding_window=24, suppress=sup, max_new_tokens=6)
    total = 0
    for i in range(10):
        px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
        prompt = synthetic.round_prompt(i, 16, 6)
        ro = oracle.step(prompt, px.to(torch.bfloat16), grid)
        rg = engine.step(prompt, px.to("cuda"), grid)
        assert rg["ids"] == ro["ids"], f"seed {seed}: stream forked in chunk {i}"
        total += len(ro["new_tokens"])
    assert total == 70 and engine.cache.compactions >= 6
    record("unforced_stream", seed=seed, identical_tokens=total)
