"""GPU: the reference-facing surface of `StreamingCache` driven the way the reference's own code drives its cache
(`generate/streaming_cache.py:6-74`, `inference/inference.py:50-68, 106-107`), against the oracle's list-of-tensors cache:

* `prune_id_and_kv_cache` (E3) through a real arena;
* the reference's literal pruning code -- `for i, (k, v) in enumerate(cache): cache.key_cache[i] = index_select(...)` -- on
  the arena (the `key_cache[i] = t` route), followed by attention over the result;
* HF-style `update()` with and without known positions; iteration, `len`, `get_seq_length`, `update_position_ids`.

K_raw / V / ids are compared bit for bit; K_rot against the oracle's mRoPE of the arena's K_raw at the oracle's positions.
"""
import numpy as np
import pytest
import torch

from oracle import attention as o_attn
from oracle import eviction as o_ev
from oracle import positions as o_pos
from oracle.cache import ListCache
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import TINY

pytestmark = pytest.mark.gpu

GEO = TINY


def bits(t):
    return t.contiguous().view(torch.int16)


def _ids(n_rounds=3, n_vis=16):
    ids = []
    for i in range(n_rounds):
        p = synthetic.round_prompt(i, n_vis, 6)
        ids += (p if not ids or ids[-1] != 198 else p[1:]) + synthetic._text_tokens(5, 40 + i) + [151645]
    return ids


def _filled(ids, seed=0):
    """(arena cache, oracle ListCache) holding the same random K_raw / V rows for len(ids) tokens, appended through `update()`."""
    from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache
    n = len(ids)
    g = torch.Generator().manual_seed(seed)
    cache = StreamingCache(GEO, capacity=512, device="cuda")
    ref = ListCache()
    grids = [(1, 8, 8)] * 8
    cache.set_video_grids(grids)
    cache.compute_positions(ids)
    for li in range(GEO.num_layers):
        k = (torch.randn(1, GEO.num_kv_heads, n, 128, generator=g) * 0.5).to(torch.bfloat16)
        v = (torch.randn(1, GEO.num_kv_heads, n, 128, generator=g) * 0.5).to(torch.bfloat16)
        ref.update(k, v, li)
        k_all, v_all = cache.update(k.cuda(), v.cuda(), li)
        assert k_all.shape == (1, GEO.num_kv_heads, n, 128) and torch.equal(bits(k_all.cpu()), bits(k))      # raw keys come back
    return cache, ref, grids


def _expect_rot(k_raw_cpu, ids, grids):
    """Oracle K_rot [H, n, D] of K_raw [H, n, D] at the contiguous positions of `ids`."""
    pos, _ = o_pos.get_rope_index(ids, grids, GEO.spatial_merge_size, GEO.tokens_per_second)
    tc, ts = o_pos.rotary_table(512, o_pos.default_inv_freq(128, GEO.rope_theta))
    idx = pos[:, 0].long()
    cos, sin = torch.cat((tc[idx], tc[idx]), dim=-1)[:, None], torch.cat((ts[idx], ts[idx]), dim=-1)[:, None]
    _, rot = o_pos.apply_mrope(k_raw_cpu[None, :1, :1], k_raw_cpu[None], cos, sin, GEO.mrope_section)
    return rot[0], pos


def _check_against(cache, ref, ids, grids, what):
    n = len(ids)
    assert cache.get_seq_length() == n == ref.get_seq_length(), what
    for li in range(GEO.num_layers):
        k_raw, k_rot, v = (t.cpu() for t in cache.planes(li))
        assert torch.equal(bits(k_raw), bits(ref.key_cache[li][0])), f"{what}: K_raw layer {li}"
        assert torch.equal(bits(v), bits(ref.value_cache[li][0])), f"{what}: V layer {li}"
        want, _ = _expect_rot(k_raw, ids, grids)
        assert torch.equal(bits(k_rot), bits(want)), f"{what}: K_rot layer {li}"


def test_update_appends_and_rotates_when_positions_are_known():
    ids = _ids()
    cache, ref, grids = _filled(ids)
    assert not cache._rot_stale
    _check_against(cache, ref, ids, grids, "update")
    # iteration / len / key_cache views: the reference reads `for i, (k_layer, v_layer) in enumerate(past_key_values)`
    assert len(cache) == GEO.num_layers
    for li, (k, v) in enumerate(cache):
        assert k.shape == (1, GEO.num_kv_heads, len(ids), 128)
        assert torch.equal(bits(k.cpu()), bits(ref.key_cache[li])) and torch.equal(bits(v.cpu()), bits(ref.value_cache[li]))
        assert torch.equal(bits(cache.key_cache[li].cpu()), bits(ref.key_cache[li]))
    # update_position_ids stores and returns what it was given (reference :10-28)
    p = torch.arange(6.0).reshape(3, 1, 2)
    assert cache.update_position_ids(p) is p and cache.update_position_ids(None) is p


def test_prune_id_and_kv_cache_through_the_arena():
    """E3 (`inference.py:50-61`): closed-interval delete from ids and from every layer, several intervals in a row."""
    from streaming_vlm_b200.inference.inference import contiguous_id_and_kv, prune_id_and_kv_cache
    ids = _ids()
    cache, ref, grids = _filled(ids, seed=1)
    ids_t = torch.tensor([ids], device="cuda")
    ids_o = torch.tensor([ids])
    def block(cur, k):       # k-th <|vision_start|> ... <|vision_end|> span of the current ids (a whole block: run and grid stay consistent)
        s_ = [i for i, t in enumerate(cur) if t == 151652][k]
        return s_, cur.index(151653, s_)

    for pick in (lambda cur: (12, 20), lambda cur: (0, 3), lambda cur: block(cur, 1), lambda cur: (len(cur) - 9, len(cur) - 4)):
        a, b = pick(ids_o[0].tolist())
        before = cache.compactions
        ids_t, cache = prune_id_and_kv_cache(ids_t, cache, a, b)
        ids_t, cache = contiguous_id_and_kv(ids_t, cache)
        ids_o, ref = o_ev.prune_id_and_kv_cache(ids_o, ref, a, b)
        assert cache.compactions == before + 1
        assert ids_t.is_cuda and ids_t[0].tolist() == ids_o[0].tolist()
        n_vid = int((ids_o[0] == 151652).sum())
        _check_against(cache, ref, ids_o[0].tolist(), grids[:max(n_vid, 1)], f"prune [{a},{b}]")


def test_reference_style_index_select_assignment_is_rerotated():
    """The reference's own pruning code (`inference.py:52-60`) run verbatim against the arena cache: `key_cache[i] = ...` leaves
    K_rot stale; the next position computation re-derives it, and attention over the result matches the oracle."""
    ids = _ids()
    cache, ref, grids = _filled(ids, seed=2)
    start_index, end_index = 12, 22               # plain-text rows of the "previous text" block
    seq_len = cache.get_seq_length()
    for i, (k_layer, v_layer) in enumerate(cache):
        indices_to_keep = list(range(start_index)) + list(range(end_index + 1, seq_len))
        idx = torch.tensor(indices_to_keep, device=k_layer.device)
        cache.key_cache[i] = torch.index_select(k_layer, 2, idx)
        cache.value_cache[i] = torch.index_select(v_layer, 2, idx)
    new_ids = ids[:start_index] + ids[end_index + 1:]
    _, ref = o_ev.prune_id_and_kv_cache(torch.tensor([ids]), ref, start_index, end_index)
    assert cache._rot_stale and cache.get_seq_length() == len(new_ids)
    # no positions for the new rows yet: reading the rotated keys must fail loudly, not return stale data
    with pytest.raises(RuntimeError, match="no positions are known"):
        cache.planes(0)
    cache.compute_positions(new_ids)                   # what the next forward does first (model_forward.py:101-110)
    assert not cache._rot_stale
    _check_against(cache, ref, new_ids, grids, "index_select assignment")
    # one decode-style attention over the re-rotated cache
    from streaming_vlm_b200 import _abi
    _abi.register_torch_ops()
    n, hq, d = len(new_ids), GEO.num_heads, 128
    q = (torch.randn(1, hq, d, generator=torch.Generator().manual_seed(5)) * 0.5).to(torch.bfloat16)
    out = torch.ops.svlm.attn_causal(cache.arena, cache.cur, 1, q.cuda(), n, d ** -0.5, 2)
    k_raw, _, v = (t.cpu() for t in cache.planes(1))
    k_rot, _ = _expect_rot(k_raw, new_ids, grids)
    want = o_attn.attention(q.transpose(0, 1), k_rot, v, causal=True)
    torch.testing.assert_close(out.cpu().float().reshape(1, -1), want.float().reshape(1, -1), rtol=2e-2, atol=6e-3)


def test_update_before_positions_is_deferred_not_lost():
    """HF-style `update()` for rows whose positions are not known yet: the rows are stored, the cache refuses to serve rotated
    keys until positions exist, then serves the right ones."""
    from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache
    ids = _ids(2)
    n = len(ids)
    g = torch.Generator().manual_seed(3)
    cache, ref = StreamingCache(GEO, capacity=256, device="cuda"), ListCache()
    for li in range(GEO.num_layers):
        k = (torch.randn(1, GEO.num_kv_heads, n, 128, generator=g) * 0.5).to(torch.bfloat16)
        v = (torch.randn(1, GEO.num_kv_heads, n, 128, generator=g) * 0.5).to(torch.bfloat16)
        ref.update(k, v, li)
        cache.update(k.cuda(), v.cuda(), li)
    assert cache._rot_stale and cache._seen_tokens == n
    with pytest.raises(RuntimeError, match="no positions are known"):
        cache.ensure_rotated()
    grids = [(1, 8, 8)] * 4
    cache.set_video_grids(grids)
    cache.compute_positions(ids)
    _check_against(cache, ref, ids, grids, "deferred update")
    # capacity is enforced on this route too
    with pytest.raises(RuntimeError, match="capacity"):
        cache.update(torch.zeros(1, GEO.num_kv_heads, 300, 128, device="cuda", dtype=torch.bfloat16),
                     torch.zeros(1, GEO.num_kv_heads, 300, 128, device="cuda", dtype=torch.bfloat16), 0)


def test_position_kernel_status_is_raised_on_the_host():
    """A run of video placeholders that disagrees with its grid used to be a status word nobody read."""
    from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache
    cache = StreamingCache(GEO, capacity=256, device="cuda")
    ids = synthetic.round_prompt(0, 16, 6)
    cache.set_video_grids([(1, 8, 12)])                # 24 merged tokens expected, 16 present
    cache.compute_positions(ids)
    with pytest.raises(ValueError, match="does not match its"):
        cache.check_status()
    cache.set_video_grids([(1, 8, 8)])
    cache.compute_positions(ids)
    cache.check_status()
    # append mode: the rotary table bounds the positions that can be handed out
    small = StreamingCache(GEO, capacity=256, device="cuda", max_position=300, pos_mode="append")
    small.reserve_positions(200)
    with pytest.raises(RuntimeError, match="rotary table"):
        small.reserve_positions(200)
