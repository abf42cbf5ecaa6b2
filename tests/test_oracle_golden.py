"""CPU: pin the oracle against fixtures produced by executing the reference's own
functions (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import eviction, positions, tokens
from oracle.cache import ListCache
from tests.stream_sim import replay, rle, sha


def _tagged(n, layers=2):
    c = ListCache()
    for li in range(layers):
        c.key_cache.append(torch.arange(n, dtype=torch.int64).view(1, 1, n, 1).clone())
        c.value_cache.append(torch.arange(n, dtype=torch.int64).view(1, 1, n, 1) + 1000000)
    return c


@pytest.fixture(scope="module")
def streams(golden_dir):
    return json.load(open(os.path.join(golden_dir, "eviction_streams.json")))


def test_get_qwen_range_matches_reference(golden_dir, streams):
    probes = json.load(open(os.path.join(golden_dir, "qwen_range_probes.json")))
    assert len(probes) > 500
    by_round = {(si, r["i"]): r["after_ids"] for si, item in enumerate(streams) for r in item["rounds"] if "after_ids" in r}
    for p in probes:
        ids = p["ids"] if "adv" in p else by_round[(p["stream"], p["round"])]
        try:
            got = list(tokens.get_qwen_range(ids, p["label"], p["index"], p["contain_lf"]))
        except IndexError:
            got = None
        assert got == p["range"], p


@pytest.mark.parametrize("si", range(5))
def test_process_past_kv_matches_reference(streams, si):
    item = streams[si]
    cfg = item["config"]
    n = 0
    for rec, before, after, gather in replay(eviction.process_past_kv, _tagged,
                                             lambda c: c.key_cache[0][0, 0, :, 0].tolist(), cfg, item["rounds"]):
        assert len(before) == rec["before_len"] and sha(before) == rec["before_sha"]
        assert len(after) == rec["after_len"], (cfg["name"], rec["i"])
        assert sha(after) == rec["after_ids_sha"], (cfg["name"], rec["i"])
        assert rle(gather) == rec["cache_runs"], (cfg["name"], rec["i"])
        if "after_ids" in rec:
            assert after == rec["after_ids"]
        n += 1
    assert n == cfg["rounds"] - 1


def test_steady_state_is_five_runs(streams):
    """defaults 16/16/512/512 settle into a 5-run gather."""
    last = streams[0]["rounds"][-1]
    assert len(last["cache_runs"]) == 5
    assert last["cache_runs"][0] == [0, 527]


def test_get_rope_index_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "rope_index.npz"))
    names = sorted({k.split("__")[0] for k in z.files})
    assert len(names) >= 10
    for nme in names:
        ids, grid, ref = z[nme + "__ids"], z[nme + "__grid"], z[nme + "__pos"]
        pos, delta = positions.get_rope_index(ids, [grid.tolist()] * 64)
        assert pos.dtype == torch.float32 and tuple(pos.shape) == ref.shape
        assert np.array_equal(pos.numpy(), ref), nme
        assert delta == int(z[nme + "__delta"].reshape(-1)[0]), nme


def _bf16_close(a, b, what):
    """cos/sin tables: fp32 `cos` of the same argument may differ in the last fp32 bit between torch's vectorised and
    scalar CPU paths (chosen by tensor shape / thread split), which flips a bf16 rounding once in ~1e5 values.  Everything
    downstream of the tables is compared bit-exactly with the tables as INPUTS."""
    ai, bi = a.contiguous().view(torch.int16).int(), b.contiguous().view(torch.int16).int()
    diff = (ai - bi).abs()
    assert int(diff.max()) <= 1, what
    assert float((diff != 0).float().mean()) < 1e-4, what


def test_mrope_apply_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "mrope_apply.npz"))
    bf = lambda a: torch.from_numpy(a).view(torch.bfloat16)
    pos = torch.from_numpy(z["pos"])
    cos, sin = positions.rotary_cos_sin(pos, positions.default_inv_freq(128, 1e6))
    _bf16_close(cos, bf(z["cos"]), "cos")
    _bf16_close(sin, bf(z["sin"]), "sin")
    # the rotation itself (section select, bf16 rounding order, right-aligned q) on the reference's own cos/sin: bit-exact
    q_rot, k_rot = positions.apply_mrope(bf(z["q"]), bf(z["k"]), bf(z["cos"]), bf(z["sin"]))
    assert torch.equal(q_rot.view(torch.int16), torch.from_numpy(z["q_rot"]))
    assert torch.equal(k_rot.view(torch.int16), torch.from_numpy(z["k_rot"]))


def test_rotary_table_equals_per_position_embedding(golden_dir):
    """The integer-position table the CUDA path indexes is the same numbers HF's rotary module produces for those
    positions."""
    z = np.load(os.path.join(golden_dir, "mrope_apply.npz"))
    pos = torch.from_numpy(z["pos"])                       # [3,1,S], integer valued
    inv = positions.default_inv_freq(128, 1e6)
    cos, sin = positions.rotary_cos_sin(pos, inv)
    tc, ts = positions.rotary_table(int(pos.max()) + 1, inv)
    for ax in range(3):
        idx = pos[ax, 0].long()
        _bf16_close(tc[idx], cos[ax, 0, :, :64], "cos table")
        _bf16_close(ts[idx], sin[ax, 0, :, 64:], "sin table")


# ------------------------------------------------------------------------------------------------ frame ingest (row f3)
def test_frame_ingest_oracle_equals_hf_processor(golden_dir):
    """oracle/frames.py == the HF Qwen2-VL video processor's own output (fixture), bit for bit in fp32."""
    import hashlib
    from oracle import frames as o_frames
    from streaming_vlm_b200 import synthetic
    from streaming_vlm_b200.config import QWEN25VL_7B
    z = np.load(os.path.join(golden_dir, "frame_ingest.npz"))
    for name in ("small", "wide"):
        fr = torch.from_numpy(z[name + "__frames_thwc"])
        got, grid = o_frames.frames_to_patches(fr)
        assert list(grid) == z[name + "__grid"][0].tolist()
        assert np.array_equal(got.numpy().view(np.int32), z[name + "__pixel_values"].view(np.int32)), name
        # the synthetic-data helper used by the benches and session tests is the same arithmetic
        mine, grid2 = synthetic.patchify(fr.permute(0, 3, 1, 2).contiguous(), QWEN25VL_7B)
        assert grid2 == grid and torch.equal(mine, got)
    seed, h, w = (int(x) for x in z["full__seed_hw"])
    fr = synthetic.frames(seed, h, w).permute(0, 2, 3, 1).contiguous()
    got, grid = o_frames.frames_to_patches(fr)
    assert grid == (1, 32, 32) and got.shape == (1024, 1176)
    assert hashlib.sha256(got.numpy().tobytes()).digest() == bytes(z["full__sha256"])


# ------------------------------------------------------------------------------------------------ Qwen2-VL variant (row f4)
def test_qwen2_vl_tower_oracle_equals_hf_module():
    """oracle/vit.py::vision_forward_qwen2 against transformers' own `Qwen2VisionTransformerPretrainedModel` (the module
    the reference patches in `inference/qwen2/`), same random weights, fp32, on CPU."""
    from transformers import Qwen2VLConfig
    from transformers.models.qwen2_vl.modeling_qwen2_vl import Qwen2VisionTransformerPretrainedModel
    from oracle import vit as o_vit
    from streaming_vlm_b200 import synthetic
    from streaming_vlm_b200.config import from_hf_config
    torch.manual_seed(0)
    cfg = Qwen2VLConfig(text_config=dict(hidden_size=256, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1, intermediate_size=256,
                                         vocab_size=152064, rope_parameters={"rope_type": "default", "rope_theta": 1e6, "mrope_section": [16, 24, 24]}),
                        vision_config=dict(depth=3, embed_dim=160, num_heads=2, mlp_ratio=2, hidden_size=256))
    geo = from_hf_config(cfg)
    assert geo.vit_kind == "qwen2" and geo.vit_hidden == 160 and geo.tokens_per_second == 1
    tower = Qwen2VisionTransformerPretrainedModel._from_config(cfg.vision_config, attn_implementation="eager").float().eval()
    w = {"model.visual." + k: v.detach().clone() for k, v in tower.state_dict().items()}
    px, grid = synthetic.patchify(synthetic.frames(2, 112, 84, n=4), geo)               # two frame pairs, non-square
    with torch.no_grad():
        out = tower(px, torch.tensor([list(grid)]))
    want = out.pooler_output if hasattr(out, "pooler_output") else out
    got = o_vit.vision_forward(px, [grid], w, geo.as_dict())
    assert got.shape == want.shape == (grid[0] * grid[1] * grid[2] // 4, 256)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_spatial_resize_oracle_vs_torchvision(golden_dir):
    """oracle/frames.py::resize_bicubic_aa against torchvision's `resize(BICUBIC, antialias=True)` on uint8 clips (fixtures):
    at most 1 LSB apart, on at most 0.1 % of the values (fp32 accumulation order at rounding boundaries); the product's tap
    builder equals the oracle's bit for bit; `smart_resize` equals the copy that ships in transformers."""
    from oracle import frames as o_frames
    from streaming_vlm_b200.utils import frames as p_frames
    z = np.load(os.path.join(golden_dir, "frame_ingest.npz"))
    for name in ("down", "mixed", "up"):
        clip, want = torch.from_numpy(z[f"resize_{name}__in_thwc"]), torch.from_numpy(z[f"resize_{name}__out_thwc"])
        got = o_frames.resize_bicubic_aa(clip, want.shape[1], want.shape[2])
        d = (got.int() - want.int()).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3, (name, int(d.max()), float((d > 0).float().mean()))
    for n_in, n_out in [(160, 112), (90, 56), (100, 140), (53, 56), (448, 448), (1280, 700), (37, 28)]:
        a, b = o_frames.aa_bicubic_taps(n_in, n_out), p_frames.aa_bicubic_taps(n_in, n_out)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.int32), b[2].view(np.int32))
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize as hf_smart_resize
    rng = np.random.default_rng(0)
    for _ in range(500):
        h, w = int(rng.integers(28, 2200)), int(rng.integers(28, 3900))
        mn, mx = int(rng.choice([56 * 56, 128 * 28 * 28])), int(rng.choice([768 * 28 * 28, 602112, 16384 * 28 * 28]))
        want = hf_smart_resize(h, w, 28, mn, mx)
        assert o_frames.smart_resize(h, w, 28, mn, mx) == want and p_frames.smart_resize(h, w, 28, mn, mx) == want


def test_append_mode_positions_equal_reference_model_forward(golden_dir):
    """`pos_mode="append"`: oracle/positions.py::append_mode_positions against the positions the reference's own `model_forward`
    (exec'd with a stub model, tests/golden/make_golden.py::make_append_positions) handed to the language model, over ten chunk
    prefills with different grids and the decode steps after them."""
    z = json.load(open(os.path.join(golden_dir, "append_positions.json")))
    state = {"last_cache_position": -1}
    cached = 0
    for i, ch in enumerate(z["chunks"]):
        for step, want in enumerate(ch["pos"]):
            new = ch["new_ids"] if step == 0 else [1000 + step]
            got = positions.append_mode_positions(new, [tuple(ch["grid"])], cached, state)
            assert got[:, 0, :].to(torch.int64).tolist() == want, f"chunk {i} step {step}"
            cached += len(new)
        assert state["last_cache_position"] == ch["last_cache_position"], f"chunk {i}"
