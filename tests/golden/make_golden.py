#!/usr/bin/env python
"""Generate the golden fixtures by EXECUTING THE REFERENCE'S OWN FUNCTIONS.

Runs only in the build container (it reads /root/reference/src, which does not exist on
the GPU box); the fixtures it writes next to this file are committed and are what the
tests read.  Nothing is copied from the reference: its functions are loaded from their
source files at run time, called on seeded synthetic inputs, and only their OUTPUTS are
stored.

    get_qwen_range                    imported as-is (utils/get_qwen_range.py)
    prune_id_and_kv_cache,
    contiguous_id_and_kv,
    process_past_kv                   exec'd from the AST of inference/inference.py (the module
                                      itself imports decord / qwen_vl_utils, absent here)
    get_rope_index                    exec'd from qwen2_5/pos_emb.py with FPS = 2.0 injected
                                      (qwen_vl_utils.vision_process.FPS of the pinned 0.0.11)
    apply_multimodal_rotary_pos_emb   exec'd from qwen2_5/language_forward.py with HF's
                                      `rotate_half` (same function in 4.52 and 5.5)
    Qwen2VLVideoProcessor             HF 5.5.0, the processor class behind the reference's `processor(videos=...)`

Usage:  python tests/golden/make_golden.py
"""
import ast
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from streaming_vlm.utils.get_qwen_range import TOKEN_IDS, get_qwen_range  # noqa: E402  (reference)
from streaming_vlm_b200 import synthetic  # noqa: E402


def _exec_functions(path, names, ns):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names), (path, names)
    exec(compile(ast.Module(body, []), path, "exec"), ns)
    return ns


class FakeCache:
    """HF 4.52 DynamicCache surface used by inference.py: key_cache/value_cache lists,
    iteration, get_seq_length.  Tensors are [1,1,S,1] holding arange(S), so that the
    surviving values ARE the gather map."""

    def __init__(self, n, layers=2):
        self.key_cache = [torch.arange(n, dtype=torch.int64).view(1, 1, n, 1).clone() for _ in range(layers)]
        self.value_cache = [torch.arange(n, dtype=torch.int64).view(1, 1, n, 1).clone() + 1000000 for _ in range(layers)]

    def __iter__(self):
        return iter(zip(self.key_cache, self.value_cache))

    def get_seq_length(self):
        return self.key_cache[0].shape[2]


def rle(a):
    """[(src_start, length), ...] of maximal +1 runs."""
    runs = []
    for x in a:
        if runs and runs[-1][0] + runs[-1][1] == x:
            runs[-1][1] += 1
        else:
            runs.append([int(x), 1])
    return runs


def sha(ids):
    return hashlib.sha256(np.asarray(ids, dtype=np.int64).tobytes()).hexdigest()[:16]


def simulate(ns, cfg):
    """Drive process_past_kv exactly as streaming_inference does (inference.py:309-489) on a
    synthetic stream; generated text is a seeded list of plain ids of variable length."""
    rng = np.random.RandomState(cfg["seed"])
    hist, clips, pixels = [], [], []
    prev, cache = None, None
    rounds = []
    for i in range(cfg["rounds"]):
        before = None if prev is None else prev.flatten().tolist()
        if cache is not None:
            cache = FakeCache(cache.get_seq_length())
        cache, prev, clips, pixels = ns["process_past_kv"](
            cache, i, text_round=cfg["text_round"], visual_round=cfg["window"], full_conversation_history=hist,
            prev_generated_ids=prev, assistant_start_bias=3, assistant_end_bias=2, recent_video_window_clips=clips,
            recent_pixel_values_videos=pixels, text_sink=cfg["text_sink"], text_sliding_window=cfg["text_sliding_window"])
        rec = {"i": i}
        if before is not None:
            g = cache.key_cache[0][0, 0, :, 0].tolist()
            assert cache.value_cache[1][0, 0, :, 0].tolist() == [x + 1000000 for x in g]
            rec.update(before_len=len(before), before_sha=sha(before), after_ids_sha=sha(prev.flatten().tolist()),
                       after_len=int(prev.shape[1]), cache_runs=rle(g))
            if i in cfg.get("dump_rounds", ()):
                rec["after_ids"] = prev.flatten().tolist()
        rounds.append(rec)
        clips.append(None)
        prompt = synthetic.round_prompt(i, cfg["n_vis"], cfg["prev_text"], 8, cfg["seed"])
        if i == 0:
            hist[:] = [{"role": "previous text", "content": "p" * 10},
                       {"role": "user", "content": [{"type": "text", "text": "T"}, {"type": "video"}, {"type": "text", "text": "q"}]}]
        else:
            hist.append({"role": "user", "content": [{"type": "text", "text": "T"}, {"type": "video"}]})
        ids = prompt if prev is None else (prev.flatten().tolist() + (prompt if prev[0, -1].item() != 198 else prompt[1:]))
        pixels.append(None)
        input_len = len(ids)
        n_new = int(rng.randint(cfg["gen_lo"], cfg["gen_hi"] + 1))
        eos = bool(rng.rand() < cfg["p_eos"])
        gen = rng.randint(1000, 100000, size=n_new).tolist()
        gen = [g if g not in (872, 77091, 1462, 198, 19702, 1467) else 1000 for g in gen]
        if eos:
            gen[-1] = 151645
        seq = ids + gen
        cache_len = len(seq) - 1                   # the last sampled token is never fed back
        if seq[-1] != 151645:
            seq = seq + [151645]                   # inference.py:457-459
        prev = torch.tensor([seq], dtype=torch.long)
        cache = FakeCache(cache_len)
        hist.append({"role": "assistant", "content": "x" * 8 + " ..."})
        rec["input_len"] = input_len
        rec["gen"] = gen
    return rounds


def main():
    ns = {"torch": torch, "get_qwen_range": get_qwen_range, "TOKEN_IDS": TOKEN_IDS, "breakpoint": lambda: (_ for _ in ()).throw(RuntimeError("breakpoint"))}
    _exec_functions(f"{REF}/streaming_vlm/inference/inference.py",
                    ["prune_id_and_kv_cache", "contiguous_id_and_kv", "process_past_kv"], ns)

    # ---------------------------------------------------------------- eviction streams
    configs = [
        dict(name="default_16_16_512_512", window=16, text_round=16, text_sink=512, text_sliding_window=512, rounds=150,
             n_vis=256, prev_text=24, gen_lo=4, gen_hi=20, p_eos=0.5, seed=1, dump_rounds=[16, 17, 149]),
        dict(name="small_4_4_16_24", window=4, text_round=4, text_sink=16, text_sliding_window=24, rounds=40,
             n_vis=16, prev_text=6, gen_lo=3, gen_hi=9, p_eos=0.5, seed=2, dump_rounds=[4, 5, 20, 39]),
        dict(name="vis_gt_text_6_3_8_8", window=6, text_round=3, text_sink=8, text_sliding_window=8, rounds=30,
             n_vis=16, prev_text=4, gen_lo=3, gen_hi=9, p_eos=0.3, seed=3, dump_rounds=[6, 29]),
        dict(name="text_gt_vis_3_6_8_8", window=3, text_round=6, text_sink=8, text_sliding_window=8, rounds=30,
             n_vis=16, prev_text=4, gen_lo=3, gen_hi=9, p_eos=0.3, seed=4, dump_rounds=[6, 29]),
        dict(name="no_trim_4_4_none", window=4, text_round=4, text_sink=None, text_sliding_window=None, rounds=20,
             n_vis=16, prev_text=4, gen_lo=3, gen_hi=6, p_eos=1.0, seed=5, dump_rounds=[19]),
    ]
    out = []
    for cfg in configs:
        rounds = simulate(ns, cfg)
        out.append({"config": cfg, "rounds": rounds})
        last = rounds[-1]
        print(cfg["name"], "last before/after", last.get("before_len"), last.get("after_len"), "runs", last.get("cache_runs"))
    json.dump(out, open(os.path.join(HERE, "eviction_streams.json"), "w"))

    # ---------------------------------------------------------------- get_qwen_range probes
    probes = []
    for cfg_i, item in enumerate(out[:4]):
        for rec in item["rounds"]:
            if "after_ids" not in rec:
                continue
            ids = rec["after_ids"]
            for label in ["user", "previous text", "assistant", "vision", "user_text"]:
                for idx in [0, 1, -1, -2]:
                    for lf in [True, False]:
                        try:
                            r = list(get_qwen_range(torch.tensor([ids]), label, idx, contain_lf=lf))
                        except IndexError:
                            r = None
                        probes.append({"stream": cfg_i, "round": rec["i"], "label": label, "index": idx, "contain_lf": lf, "range": r})
    # adversarial: unterminated segment, adjacent patterns, lf at the very end
    adv = [
        [151644, 872, 198, 5, 6, 151645, 198, 151644, 872, 198, 7],
        [151644, 77091, 198, 151645, 151644, 77091, 198, 9, 151645, 198],
        [151652, 151656, 151656, 151653, 151652, 151653],
        [1462, 3, 4, 151652, 151656, 151653, 1462, 151652],
        [151644, 19702, 1467, 198, 151645],
        [5, 6, 7],
    ]
    for a_i, ids in enumerate(adv):
        for label in ["user", "previous text", "assistant", "vision", "user_text"]:
            for idx in [0, 1, -1]:
                for lf in [True, False]:
                    try:
                        r = list(get_qwen_range(torch.tensor([ids]), label, idx, contain_lf=lf))
                    except IndexError:
                        r = None
                    probes.append({"adv": a_i, "ids": ids, "label": label, "index": idx, "contain_lf": lf, "range": r})
    json.dump(probes, open(os.path.join(HERE, "qwen_range_probes.json"), "w"))
    print("range probes", len(probes))

    # ---------------------------------------------------------------- get_rope_index
    stub = types.ModuleType("qwen_vl_utils"); vp = types.ModuleType("qwen_vl_utils.vision_process"); vp.FPS = 2.0
    ns2 = {"torch": torch, "Optional": __import__("typing").Optional, "Tuple": __import__("typing").Tuple, "FPS": 2.0}
    _exec_functions(f"{REF}/streaming_vlm/inference/qwen2_5/pos_emb.py", ["get_rope_index"], ns2)

    class _Cfg:  # the attributes get_rope_index reads (pos_emb.py:45-48, :126)
        class vision_config:
            spatial_merge_size = 2
            tokens_per_second = 2
        image_token_id = 151655
        video_token_id = 151656
        vision_start_token_id = 151652
    self_stub = types.SimpleNamespace(config=_Cfg)
    rope = {}
    cases = []
    for item in out[:4]:
        cfg = item["config"]
        for rec in item["rounds"]:
            if "after_ids" in rec:
                side = int(round(cfg["n_vis"] ** 0.5)) * 2
                cases.append((rec["after_ids"], [1, side, side], f"{cfg['name']}_r{rec['i']}"))
    # extra: multi-frame grid (t=2), non-square grid, trailing zeros as in decode steps
    ids_t2 = synthetic.SYSTEM_PROMPT + [151644, 872, 198, 1462, 5, 151652] + [151656] * (2 * 2 * 3) + [151653, 151645, 198, 7, 8, 0, 0, 0]
    cases.append((ids_t2, [2, 4, 6], "t2_h2_w3_trailing_zeros"))
    for ids, grid, name in cases:
        t = torch.tensor([ids], dtype=torch.long)
        nvid = int((t == 151656).sum().item() // max(1, (grid[0] * (grid[1] // 2) * (grid[2] // 2))))
        g = torch.tensor([grid] * (nvid + 3), dtype=torch.long)   # grid list longer than needed, as in the reference (:415)
        pos, delta = ns2["get_rope_index"](self_stub, t, None, g, None, torch.ones_like(t, dtype=torch.bool))
        rope[name + "__ids"] = np.asarray(ids, dtype=np.int64)
        rope[name + "__grid"] = np.asarray(grid, dtype=np.int64)
        rope[name + "__pos"] = pos.numpy()
        rope[name + "__delta"] = np.asarray(delta.numpy())
        print("rope", name, pos.shape, float(pos.max()))
    np.savez_compressed(os.path.join(HERE, "rope_index.npz"), **rope)

    # ---------------------------------------------------------------- mRoPE apply (bf16 op order)
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VLRotaryEmbedding, rotate_half
    ns3 = {"torch": torch, "rotate_half": rotate_half}
    _exec_functions(f"{REF}/streaming_vlm/inference/qwen2_5/language_forward.py", ["apply_multimodal_rotary_pos_emb"], ns3)
    g = torch.Generator().manual_seed(11)
    name = "default_16_16_512_512_r17"
    pos = torch.from_numpy(rope[name + "__pos"])[:, :, :700]
    inv_freq = 1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).to(torch.float32) / 128))
    # HF rotary forward (modeling_qwen2_5_vl.py:595-608), inlined because the module wants a config object
    freqs = (inv_freq[None, None, :, None].float().expand(3, 1, -1, 1) @ pos[:, :, None, :].float()).transpose(2, 3)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
    q = torch.randn(1, 4, 37, 128, generator=g).to(torch.bfloat16)
    k = torch.randn(1, 2, 700, 128, generator=g).to(torch.bfloat16)
    qe, ke = ns3["apply_multimodal_rotary_pos_emb"](q, k, cos, sin, [16, 24, 24])
    np.savez_compressed(os.path.join(HERE, "mrope_apply.npz"), pos=pos.numpy(),
                        q=q.view(torch.int16).numpy(), k=k.view(torch.int16).numpy(),
                        cos=cos.view(torch.int16).numpy(), sin=sin.view(torch.int16).numpy(),
                        q_rot=qe.view(torch.int16).numpy(), k_rot=ke.view(torch.int16).numpy())
    print("mrope", qe.shape, ke.shape)
    make_frame_ingest()
    make_append_positions(ns2["get_rope_index"])


def make_append_positions(ref_get_rope_index):
    """`pos_mode="append"`: the reference's `model_forward` (qwen2_5/model_forward.py:30-176) itself, exec'd from its source and
    called with a stub `self` whose `get_rope_index` is the reference's and whose `language_model` records the `position_ids`
    it is handed.  `inputs_embeds` is passed in, so no weights are involved: what is recorded is exactly the position policy
    (:112-153) over a stream of chunk prefills (new ids only, as HF's generate passes them) and single-token decode steps."""
    import typing
    ns = {"torch": torch, "Optional": typing.Optional, "List": typing.List, "Tuple": typing.Tuple, "Union": typing.Union,
          "StreamingArgs": object, "Qwen2_5_VLModelOutputWithPast": lambda **kw: types.SimpleNamespace(**kw)}
    _exec_functions(f"{REF}/streaming_vlm/inference/qwen2_5/model_forward.py", ["get_1d_rope_index", "model_forward"], ns)
    seen = []

    class Self:
        rope_deltas = None
        config = types.SimpleNamespace(output_attentions=False, output_hidden_states=False, use_return_dict=True,
                                       vision_config=types.SimpleNamespace(spatial_merge_size=2, tokens_per_second=2),
                                       image_token_id=151655, video_token_id=151656, vision_start_token_id=151652)

        def get_rope_index(self, *a, **k):
            return ref_get_rope_index(self, *a, **k)

        def language_model(self, **kw):
            seen.append(kw["position_ids"].clone())
            return types.SimpleNamespace(last_hidden_state=None, past_key_values=kw["past_key_values"], hidden_states=None, attentions=None)

    me = Self()
    args = types.SimpleNamespace(pos_mode="append", all_text=False, input_ids=None, video_grid_thw=None, second_per_grid_ts=None,
                                 last_cache_position=-1)
    rng = np.random.RandomState(7)
    cached, chunks = 0, []
    for i in range(10):
        n_vis = [16, 64, 16, 256][i % 4]
        side = int(round((n_vis * 4) ** 0.5))
        grid = torch.tensor([[1, side, side]])
        new = ([151645] if i else []) + synthetic.round_prompt(i, n_vis, 6, 8, 3)
        n_dec = int(rng.randint(1, 7))
        rec = {"new_ids": new, "grid": grid[0].tolist(), "n_decode": n_dec, "pos": []}
        for step in range(1 + n_dec):
            ids = torch.tensor([new]) if step == 0 else torch.tensor([[1000 + step]])
            t = ids.shape[1]
            cache = types.SimpleNamespace(get_seq_length=lambda c=cached: c)
            ns["model_forward"](me, input_ids=ids, attention_mask=torch.ones(1, cached + t, dtype=torch.long), past_key_values=cache,
                                inputs_embeds=torch.zeros(1, t, 8), video_grid_thw=grid if step == 0 else None,
                                cache_position=torch.arange(cached, cached + t), streaming_args=args)
            rec["pos"].append(seen[-1][:, 0, :].to(torch.int64).tolist())
            cached += t
        rec["last_cache_position"] = int(args.last_cache_position)
        chunks.append(rec)
        cached = max(cached - int(rng.randint(0, 40)), 1)          # an eviction in between: irrelevant to append-mode positions
    json.dump({"chunks": chunks}, open(os.path.join(HERE, "append_positions.json"), "w"))
    print("append positions", len(chunks), "chunks, last position", chunks[-1]["last_cache_position"])


def make_frame_ingest():
    """Frame ingest: the HF processor the reference calls (`processor(text=..., videos=clip)`, inference.py:390-395) run on
    seeded uint8 frames that are already at their target size.  transformers 5.5.0 is the only version in the image (the
    reference pins 4.52.4, use_fast=False); stored: the full fp32 output for a 112 x 112 clip and a SHA-256 for 448 x 448."""
    from transformers import Qwen2VLVideoProcessor
    vp = Qwen2VLVideoProcessor(min_pixels=4 * 28 * 28, max_pixels=16384 * 28 * 28)
    out = {}
    for name, (h, w, seed) in {"small": (112, 112, 3), "wide": (56, 168, 4)}.items():
        fr = synthetic.frames(seed, h, w)                                  # uint8 [2, 3, H, W]
        r = vp(videos=[fr], return_tensors="pt", do_resize=False)
        out[name + "__frames_thwc"] = fr.permute(0, 2, 3, 1).contiguous().numpy()
        out[name + "__pixel_values"] = r["pixel_values_videos"].numpy()
        out[name + "__grid"] = r["video_grid_thw"].numpy()
    # spatial resize exactly as livecc_utils' _spatial_resize_video calls it (video_process_patch.py:147-152), on uint8 TCHW clips
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF
    for name, (h, w, h2, w2, seed) in {"down": (90, 160, 56, 112, 11), "mixed": (37, 53, 28, 56, 12), "up": (50, 50, 84, 84, 13)}.items():
        clip = synthetic.frames(seed, h, w)                                # uint8 [2, 3, h, w]
        res = TF.resize(clip, [h2, w2], interpolation=InterpolationMode.BICUBIC, antialias=True)
        out["resize_" + name + "__in_thwc"] = clip.permute(0, 2, 3, 1).contiguous().numpy()
        out["resize_" + name + "__out_thwc"] = res.permute(0, 2, 3, 1).contiguous().numpy()
    fr = synthetic.frames(5, 448, 448)
    r = vp(videos=[fr], return_tensors="pt", do_resize=False)
    out["full__sha256"] = np.frombuffer(hashlib.sha256(r["pixel_values_videos"].numpy().tobytes()).digest(), dtype=np.uint8)
    out["full__seed_hw"] = np.array([5, 448, 448])
    np.savez_compressed(os.path.join(HERE, "frame_ingest.npz"), **out)
    print("frame ingest", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
