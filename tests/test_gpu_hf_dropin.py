"""GPU: the drop-in boundary.  A stock transformers Qwen2_5_VLForConditionalGeneration (tiny, random init) is
patched with `convert_qwen2_5_to_streaming` and driven exactly like the reference's session loop drives it
(`inference/inference.py:319, 397-416, 440-451`): `process_past_kv` -> `model.generate(**inputs,
past_key_values=..., streaming_args=...)`.  Checked against the CPU oracle run on the same weights, for

* the FAST path (the default: `generate` runs the turn on the CUDA-graph session runner built over the model's own parameters),
* the EAGER path (`SVLM_DROPIN_EAGER=1`: every patched HF forward executes per step),
* both position modes of `StreamingArgs` on both paths, and the Qwen2-VL variant (eager).

The stream is teacher-forced with the oracle's tokens (`forced_tokens`, the purpose of the reference's --gt_json mode), so a
near-tie cannot fork it; every selected token is compared and the number of differences is recorded and bounded.
"""
import pytest
import torch

from oracle.model import Session
from streaming_vlm_b200 import synthetic
from streaming_vlm_b200.config import from_hf_config
from tests.parity_log import MAX_NEAR_TIE_FRACTION, near_tie, record

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.08


def _tiny_model(family):
    text = dict(hidden_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=1024,
                vocab_size=152064, rms_norm_eps=1e-6, rope_parameters={"rope_type": "default", "rope_theta": 1e6, "mrope_section": [16, 24, 24]})
    if family == "qwen2_5":
        from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
        from streaming_vlm_b200.inference.qwen2_5.patch_model import convert_qwen2_5_to_streaming as convert
        cfg = Qwen2_5_VLConfig(text_config=text, vision_config=dict(depth=4, hidden_size=320, num_heads=4, intermediate_size=640, out_hidden_size=512,
                                                                    fullatt_block_indexes=[1, 3], tokens_per_second=2))
        return cfg, Qwen2_5_VLForConditionalGeneration(cfg), convert
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    from streaming_vlm_b200.inference.qwen2.patch_model import convert_qwen2_to_streaming as convert
    cfg = Qwen2VLConfig(text_config=text, vision_config=dict(depth=3, embed_dim=320, num_heads=4, mlp_ratio=2, hidden_size=512))
    return cfg, Qwen2VLForConditionalGeneration(cfg), convert


def _drive(family, pos_mode, fast, monkeypatch, n_chunks=8, repetition_penalty=1.0, pass_cache=True):
    """The reference's loop body (product side) next to the oracle session; returns (mismatches, tokens compared, model, cache)."""
    from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache
    from streaming_vlm_b200.inference.inference import process_past_kv
    from streaming_vlm_b200.inference.streaming_args import StreamingArgs
    monkeypatch.setenv("SVLM_DROPIN_EAGER", "0" if fast else "1")
    torch.manual_seed(0)
    cfg, model, convert = _tiny_model(family)
    model = model.to(torch.bfloat16).eval()
    w_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    geo = from_hf_config(cfg)
    assert geo.num_layers == 2 and geo.vit_kind == family and geo.tokens_per_second == (2 if family == "qwen2_5" else 1)
    model = convert(model.to("cuda"))
    sup = synthetic.suppress_mask(geo.vocab)
    win, tr, sink, slide = 2, 2, 8, 8
    oracle = Session(w_cpu, geo.as_dict(), win, tr, sink, slide, suppress=sup, max_new_tokens=5, pos_mode=pos_mode,
                     repetition_penalty=repetition_penalty)
    args = StreamingArgs(pos_mode=pos_mode)
    cache = StreamingCache(geo, capacity=1024, device="cuda", max_position=4096) if pass_cache else None
    history, clips, pixels = [], [], []
    prev, past = None, None
    mism = total = 0
    for i in range(n_chunks):
        px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
        prompt = synthetic.round_prompt(i, 16, 6)
        ro = oracle.step(prompt, px.to(torch.bfloat16), grid)
        # ---- the reference's loop body, product side (inference.py:319, 397-416, 440-463)
        past, prev, clips, pixels = process_past_kv(past, i, tr, win, history, prev, 3, 2, clips, pixels, sink, slide)
        clips.append(None)
        if i == 0:
            history[:] = [{"role": "previous text", "content": ""},
                          {"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}, {"type": "text", "text": "q"}]}]
        else:
            history.append({"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}]})
        ids = torch.tensor([prompt], device="cuda")
        if prev is not None:
            ids = torch.cat([prev, ids if prev[0, -1].item() != 198 else ids[:, 1:]], dim=1)
        pixels.append(None)
        g = torch.tensor([list(grid)], device="cuda")
        if pos_mode == "shrink":                     # inference.py:408-416
            args.input_ids = ids
            args.video_grid_thw = g if i == 0 else torch.cat([args.video_grid_thw, g], dim=0)
        out = model.generate(input_ids=ids, pixel_values_videos=px.to("cuda"), video_grid_thw=g, past_key_values=past if past is not None else cache,
                             max_new_tokens=5, use_cache=True, return_dict_in_generate=True, do_sample=False,
                             repetition_penalty=repetition_penalty, streaming_args=args, pad_token_id=151645,
                             suppress_tokens=sup.to("cuda"), forced_tokens=ro["new_tokens"], cache_capacity=1024)
        past = out.past_key_values
        seq = out.sequences
        if seq[0, -1].item() != 151645:
            seq = torch.cat([seq, torch.tensor([[151645]], device="cuda")], dim=1)
        prev = seq.clone()
        history.append({"role": "assistant", "content": "x ..."})
        assert seq[0].tolist() == ro["ids"], f"chunk {i}: ids"
        assert past.get_seq_length() == ro["cache_len"], f"chunk {i}: cache length"
        for j, (tg, to) in enumerate(zip(model._svlm_last_selected, ro["new_tokens"])):
            total += 1
            if tg != to:
                mism += 1
                lo = ro["step_logits"][j]
                assert near_tie(lo, tg, to), f"chunk {i} token {j}: selected {tg}, oracle {to}, oracle margin {lo[to] - lo[tg]} top {lo[to]}"
    if pos_mode == "append":
        assert args.last_cache_position == oracle.pos_state["last_cache_position"]
        assert int(past.next_pos) == oracle.pos_state["last_cache_position"] + 1
    return mism, total, model, past


@pytest.mark.parametrize("pos_mode", ["shrink", "append"])
@pytest.mark.parametrize("fast", [True, False], ids=["fast", "eager"])
def test_patched_qwen2_5_matches_oracle(fast, pos_mode, monkeypatch):
    """`convert_qwen2_5_to_streaming` (qwen2_5/patch_model.py:18-38) + the reference's loop, both executions of `generate`."""
    mism, total, model, cache = _drive("qwen2_5", pos_mode, fast, monkeypatch)
    record("hf_dropin", path="fast" if fast else "eager", pos_mode=pos_mode, mismatches=mism, tokens=total)
    assert cache.compactions >= 4
    assert (getattr(model, "_svlm_engine", None) is not None) == fast          # the fast path really ran on the session runner
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)
    if fast:
        eng = model._svlm_engine
        # one copy of the weights: the HF modules' parameters are views of the engine's fused tensors
        att = model.model.language_model.layers[0].self_attn
        assert att.q_proj.weight.data_ptr() == eng.layers[0]["wqkv"].data_ptr()
        assert att.k_proj.weight.data_ptr() == eng.layers[0]["wqkv"][att.q_proj.weight.shape[0]:].data_ptr()
        assert model.lm_head.weight.data_ptr() == eng.lm_head.data_ptr()
        assert len(eng._pf_graphs) <= eng._pf_graphs.limit and len(eng._graphs) <= eng._graphs.limit


def test_patched_qwen2_5_cache_created_by_generate(monkeypatch):
    """`past_key_values=None` on the first call (as the reference's loop starts, inference.py:297): `generate` creates the cache
    and hands it back in `.past_key_values`."""
    mism, total, model, cache = _drive("qwen2_5", "shrink", True, monkeypatch, n_chunks=5, pass_cache=False)
    record("hf_dropin_cache_created", mismatches=mism, tokens=total)
    assert cache is model._svlm_engine.cache and cache.compactions >= 2
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_patched_qwen2_5_repetition_penalty_matches_oracle(monkeypatch):
    """The reference's repetition penalty (1.05 over every id of the sequence so far) inside the lm_head kernel, greedy
    selection of the processed scores, against the oracle's restatement of HF's RepetitionPenaltyLogitsProcessor."""
    for fast in (True, False):
        mism, total, _, _ = _drive("qwen2_5", "shrink", fast, monkeypatch, n_chunks=6, repetition_penalty=1.05)
        record("hf_dropin_penalty", path="fast" if fast else "eager", mismatches=mism, tokens=total)
        assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_patched_qwen2_matches_oracle(monkeypatch):
    """`convert_qwen2_to_streaming` (reference inference/qwen2/patch_model.py): patched eager forwards."""
    mism, total, model, cache = _drive("qwen2", "shrink", True, monkeypatch, n_chunks=6)
    record("hf_dropin_qwen2", mismatches=mism, tokens=total)
    assert cache.compactions >= 3 and getattr(model, "_svlm_engine", None) is None
    assert mism <= max(1, int(total * MAX_NEAR_TIE_FRACTION)), (mism, total)


def test_reference_generation_config_runs_on_the_fast_path(monkeypatch):
    """`do_sample=True, temperature=0.9, repetition_penalty=1.05` (inference.py:41-42, 440-451) through the patched generate:
    same seed -> same stream; the stream stays inside the grammar; a different seed gives a different one."""
    from streaming_vlm_b200.inference.generate.streaming_cache import StreamingCache
    from streaming_vlm_b200.inference.inference import process_past_kv
    from streaming_vlm_b200.inference.streaming_args import StreamingArgs
    monkeypatch.setenv("SVLM_DROPIN_EAGER", "0")
    torch.manual_seed(0)
    cfg, model, convert = _tiny_model("qwen2_5")
    geo = from_hf_config(cfg)
    model = convert(model.to(torch.bfloat16).eval().to("cuda"))
    sup = synthetic.suppress_mask(geo.vocab).to("cuda")
    streams = []
    for seed in (7, 7, 8):
        args = StreamingArgs(pos_mode="shrink")
        history, clips, pixels, prev, past, toks = [], [], [], None, None, []
        cache = StreamingCache(geo, capacity=1024, device="cuda")
        for i in range(6):
            px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
            past, prev, clips, pixels = process_past_kv(past, i, 2, 2, history, prev, 3, 2, clips, pixels, 8, 8)
            clips.append(None)
            if i == 0:
                history[:] = [{"role": "previous text", "content": ""},
                              {"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}, {"type": "text", "text": "q"}]}]
            else:
                history.append({"role": "user", "content": [{"type": "text", "text": "Time"}, {"type": "video"}]})
            ids = torch.tensor([synthetic.round_prompt(i, 16, 6)], device="cuda")
            if prev is not None:
                ids = torch.cat([prev, ids if prev[0, -1].item() != 198 else ids[:, 1:]], dim=1)
            pixels.append(None)
            args.input_ids = ids
            g = torch.tensor([list(grid)], device="cuda")
            args.video_grid_thw = g if i == 0 else torch.cat([args.video_grid_thw, g], dim=0)
            out = model.generate(input_ids=ids, pixel_values_videos=px.to("cuda"), video_grid_thw=g, past_key_values=past if past is not None else cache,
                                 max_new_tokens=6, use_cache=True, return_dict_in_generate=True, do_sample=True, repetition_penalty=1.05,
                                 streaming_args=args, pad_token_id=151645, temperature=0.9, suppress_tokens=sup, seed=seed * 1000 + i)
            past, seq = out.past_key_values, out.sequences
            new = seq[0, ids.shape[1]:].tolist()
            assert 1 <= len(new) <= 6 and not any(bool(sup[t]) for t in new)
            toks.append(new)
            if seq[0, -1].item() != 151645:
                seq = torch.cat([seq, torch.tensor([[151645]], device="cuda")], dim=1)
            prev = seq.clone()
            history.append({"role": "assistant", "content": "x ..."})
        streams.append(toks)
    assert streams[0] == streams[1], "same seeds must reproduce the sampled stream"
    assert streams[0] != streams[2], "different seeds gave the same sampled stream"
