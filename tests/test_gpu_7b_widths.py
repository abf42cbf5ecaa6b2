"""GPU parity at the 7B WIDTHS (hidden 3584, 28 q / 4 kv heads -> G = 7, MLP 18944, vocab 152064; ViT 1280 wide, 16 heads of 80,
MLP 3420 padded to 3456, merger 5120 -> 3584), one steady-state chunk of BASELINE.json configs[1] (S ~ 5.8k retained tokens,
~300 new tokens, 5-run eviction) end to end against the CPU oracle.  The oracle cannot run 28 layers / 32 ViT blocks at this
width in test time, so the DEPTH is cut to 4 LLM layers and 8 ViT blocks (1 full-attention) -- every kernel runs at the
shapes, tile counts and paddings of the real model; depth only repeats them.

Compared: pruned ids and gather runs (bit-exact), position ids (bit-exact), video embeddings, per-layer attention outputs and
last-position logits (stated tolerances, measured values recorded), greedy tokens of the turn (teacher-forced, count recorded).
"""
from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle.cache import ListCache
from oracle.model import Session
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import QWEN25VL_7B as G7
from tests.parity_log import record

pytestmark = pytest.mark.gpu

LLM_LAYERS, VIT_BLOCKS, NEW_TOKENS, ROUNDS = 4, 8, 4, 200


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def test_one_steady_state_chunk_at_7b_widths():
    from streaming_vlm_b200.engine import StreamingEngine
    geo = replace(G7, num_layers=LLM_LAYERS, vit_depth=VIT_BLOCKS, fullatt_block_indexes=(VIT_BLOCKS - 1,))
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    w = synthetic.block_weights_cpu(G7, seed=0, layers=LLM_LAYERS, vit_blocks=VIT_BLOCKS)
    sup = synthetic.suppress_mask(geo.vocab)
    engine = StreamingEngine({k: v.clone() for k, v in w.items()}, geo, capacity=8192, suppress=sup, max_new_tokens=NEW_TOKENS)
    state = synthetic.fast_forward(ROUNDS, new_tokens=20)
    engine.load_state({k: (list(v) if isinstance(v, list) else v) for k, v in state.items()}, seed=3)
    n = engine.cache.seq_len
    assert 5000 < n < 6500
    # the oracle session in the same state, its list cache holding the arena's own raw keys / values
    oracle = Session(w, geo.as_dict(), max_new_tokens=NEW_TOKENS, suppress=sup)
    st2 = synthetic.fast_forward(ROUNDS, new_tokens=20)
    oracle.prev_ids = torch.tensor([st2["prev_ids"]], dtype=torch.long)
    oracle.history, oracle.clips, oracle.pixels = st2["history"], st2["clips"], st2["pixels"]
    oracle.grids, oracle.i = list(st2["grids"]), st2["i"]
    oracle.cache = ListCache()
    for li in range(LLM_LAYERS):
        k_raw, _, v = engine.cache.planes(li)
        oracle.cache.key_cache.append(k_raw.cpu()[None].clone())
        oracle.cache.value_cache.append(v.cpu()[None].clone())
    i = engine.i
    fr = synthetic.frames(i)
    px, grid = synthetic.patchify(fr, geo, dtype=torch.bfloat16)
    prompt = synthetic.round_prompt(i)
    to, tg = {}, {}
    ro = oracle.step(prompt, px, grid, taps=to)
    rg = engine.step(prompt, None, grid, taps=tg, forced=ro["new_tokens"], frames=fr.permute(0, 2, 3, 1).contiguous().cuda())
    # ---- integer state: bit-exact
    assert rg["ids"] == ro["ids"] and rg["input_len"] == ro["input_len"] and rg["cache_len"] == ro["cache_len"]
    want = [x for x in ro["gather"] if x < n]
    got = [s_ + k for s_, _, l in rg["runs"] for k in range(l)]
    assert got == want and len(rg["runs"]) == 5, rg["runs"]
    m = rg["input_len"]
    assert np.array_equal(engine.cache.pos[:, :m].cpu().numpy().astype(np.float32), to["position_ids"][:, 0, :m].numpy())
    # ---- floating point rows
    t_new = to["attn_out"][0].shape[0]
    assert 260 <= t_new <= 320 and tg["attn_out"][0].shape[0] == t_new
    ve_max, ve_rms = _rel(tg["video_embeds"], to["video_embeds"])
    att = [_rel(tg["attn_out"][li], to["attn_out"][li]) for li in range(LLM_LAYERS)]
    keep = ~sup
    lg_max, lg_rms = _rel(tg["logits"].cpu()[keep], to["logits"][keep])
    mism = sum(int(a != b) for a, b in zip(engine.last_argmax, ro["new_tokens"]))
    margins = [float(ro["step_logits"][j][b] - ro["step_logits"][j][a]) for j, (a, b) in enumerate(zip(engine.last_argmax, ro["new_tokens"])) if a != b]
    record("7b_widths_chunk", kv_len=n, new_tokens=t_new, video_embeds_max=ve_max, video_embeds_rms=ve_rms,
           attn_out_max=[a[0] for a in att], attn_out_rms=[a[1] for a in att], logits_max=lg_max, logits_rms=lg_rms,
           greedy_mismatches=mism, greedy_tokens=len(engine.last_argmax), oracle_margins_of_mismatches=margins)
    # tolerances: maximum error relative to the largest reference magnitude, and RMS error relative to the reference RMS (bf16 has 8
    # mantissa bits: one rounding = 3.9e-3 relative; 8 ViT blocks + merger, resp. up to 4 decoder layers of bf16 intermediates
    # summed in a different order).  Measured on B200: video embeddings 1.4e-2 / 1.3e-2, attention outputs <= 2.1e-2 / 1.5e-2,
    # logits 1.8e-2 / 1.8e-2 (profiles/r2a_parity_counts.jsonl).
    assert ve_max < 3e-2 and ve_rms < 2.5e-2, (ve_max, ve_rms)
    for li, (a_max, a_rms) in enumerate(att):
        assert a_max < 4e-2 and a_rms < 2.5e-2, (li, a_max, a_rms)
    assert lg_max < 4e-2 and lg_rms < 2.5e-2, (lg_max, lg_rms)
    lg_scale = float(to["logits"][keep].abs().max())
    assert all(mg < 4e-2 * lg_scale for mg in margins), margins
    assert mism <= 1, (mism, margins)
