"""GPU parity: every CUDA kernel, through the C ABI, against the CPU oracle on the same seeded inputs.

Bit-exact for integer / byte / bf16-elementwise work (compaction, positions, mRoPE, ViT RoPE);
stated bf16 tolerances for the softmax / GEMM rows.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import attention as o_attn
from oracle import positions as o_pos
from oracle import vit as o_vit

pytestmark = pytest.mark.gpu

DEV = "cuda"
SECTION = (16, 24, 24)


@pytest.fixture(scope="module")
def abi():
    from streaming_vlm_b200 import _abi
    _abi.build()
    assert _abi.lib().svlm_version() == 1
    return _abi


@pytest.fixture(scope="module")
def tables():
    cos, sin = o_pos.rotary_table(4096, o_pos.default_inv_freq(128, 1e6))
    return cos.contiguous(), sin.contiguous()


def bits(t):
    return t.contiguous().view(torch.int16)


def cos_sin_at(tables, pos_int):
    """cos/sin [3,1,n,128] for integer positions [3,n], gathered from the SAME host-built table the kernels index (HF's
    `emb = cat(freqs, freqs)` duplicates the 64 frequencies).  Recomputing cos() here instead could differ from the table in
    the last fp32 bit (vectorised vs scalar libm path) and flip a bf16 rounding once in ~1e5 values."""
    tc, ts = tables
    idx = pos_int.long()
    c, s_ = tc[idx], ts[idx]                                   # [3, n, 64]
    return torch.cat((c, c), dim=-1)[:, None], torch.cat((s_, s_), dim=-1)[:, None]


def make_arena(abi, L, H, cap, seed=0, fill=True):
    g = torch.Generator().manual_seed(seed)
    a = torch.zeros(2, 3, L, H, cap, 128, dtype=torch.bfloat16)
    if fill:
        a[0] = (torch.randn(3, L, H, cap, 128, generator=g) * 0.5).to(torch.bfloat16)
    # the TMA / UMMA tiles want a 1024-byte aligned base; torch's caching allocator only promises 512 bytes
    store = torch.zeros(a.numel() + 512, dtype=torch.bfloat16, device=DEV)
    skip = (-store.data_ptr() % 1024) // 2
    dev = store[skip:skip + a.numel()].view(a.shape)
    dev.copy_(a)
    return a, dev, abi.ArenaView(dev)


def aligned_clone(t):
    """A copy of `t` whose base is 1024-byte aligned (`clone()` only promises the caching allocator's 512 bytes)."""
    store = torch.empty(t.numel() + 512, dtype=t.dtype, device=t.device)
    skip = (-store.data_ptr() % 1024) // t.element_size()
    out = store[skip:skip + t.numel()].view(t.shape)
    out.copy_(t)
    return out


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("m,n,k,act,perm", [(256, 5120, 5120, 1, False), (256, 3584, 5120, 0, True), (100, 72, 200, 1, True),
                                            (128, 128, 64, 0, False), (1024, 3840, 1280, 0, False), (37, 512, 1280, 1, False)])
def test_gemm_bf16(abi, m, n, k, act, perm):
    g = torch.Generator().manual_seed(m + n + k)
    a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(torch.bfloat16)
    b = (torch.randn(n, generator=g) * 0.1).to(torch.bfloat16)
    ref = torch.nn.functional.linear(a.float(), w.float(), b.float()).to(torch.bfloat16)
    if act:
        ref = torch.nn.functional.gelu(ref.float()).to(torch.bfloat16)
    rows = None
    if perm:
        p = torch.randperm(m, generator=g)
        rows = p.to(torch.int32).to(DEV)
        ref_p = torch.empty_like(ref)
        ref_p[p] = ref
        ref = ref_p
    out = torch.zeros(m, n, dtype=torch.bfloat16, device=DEV)
    abi.gemm_bf16(a.to(DEV), w.to(DEV), b.to(DEV), out, act, rows)
    torch.cuda.synchronize()
    # fp32-accumulated bf16 GEMM: one bf16 ulp of the result magnitude
    torch.testing.assert_close(out.cpu().float(), ref.float(), rtol=1.6e-2, atol=1.6e-2)


def test_rmsnorm(abi):
    g = torch.Generator().manual_seed(3)
    for rows, cols in [(1024, 1280), (276, 3584), (1, 512)]:
        x = torch.randn(rows, cols, generator=g).to(torch.bfloat16)
        w = (1 + 0.1 * torch.randn(cols, generator=g)).to(torch.bfloat16)
        out = torch.empty(rows, cols, dtype=torch.bfloat16, device=DEV)
        abi.rmsnorm(x.to(DEV), w.to(DEV), out, 1e-6)
        ref = o_vit.rmsnorm(x, w, 1e-6)
        # rsqrtf vs CPU rsqrt can differ in the last fp32 bit; two bf16 roundings follow -> up to two bf16 ulps
        torch.testing.assert_close(out.cpu().float(), ref.float(), rtol=1.6e-2, atol=1e-3)


# ------------------------------------------------------------------------------------------------ positions
def test_mrope_positions_bit_exact(abi, golden_dir):
    z = np.load(os.path.join(golden_dir, "rope_index.npz"))
    names = sorted({k.split("__")[0] for k in z.files})
    for nme in names:
        ids = torch.from_numpy(z[nme + "__ids"]).to(DEV)
        grid = z[nme + "__grid"].tolist()
        n = ids.numel()
        grids = torch.tensor([grid] * 64, dtype=torch.int32, device=DEV)
        pos = torch.full((3, n + 5), -7, dtype=torch.int32, device=DEV)
        nxt = torch.zeros(1, dtype=torch.int32, device=DEV)
        status = torch.zeros(1, dtype=torch.int32, device=DEV)
        abi.mrope_positions(ids, n, grids, 64, pos, nxt, status)
        ref = z[nme + "__pos"][:, 0, :]
        assert status.item() == 0, nme
        assert np.array_equal(pos[:, :n].cpu().numpy().astype(np.float32), ref), nme
        assert nxt.item() == int(ref.max()) + 1
        assert (pos[:, n:] == -7).all()


def test_mrope_positions_long_and_edge(abi):
    """64k-token sequence (scan across all 1024 threads), empty input, grid mismatch flag."""
    from streaming_vlm_b200 import synthetic
    ids = []
    for i in range(220):
        ids += synthetic.round_prompt(i, 256, 24, 8, 0) + [5000 + i] * 9 + [151645]
    ids = ids[:65000]
    while ids[-1] == 151656:
        ids.pop()
    ref, _ = o_pos.get_rope_index(ids, [[1, 32, 32]] * 300)
    t = torch.tensor(ids, dtype=torch.int64, device=DEV)
    grids = torch.tensor([[1, 32, 32]] * 300, dtype=torch.int32, device=DEV)
    pos = torch.zeros(3, len(ids), dtype=torch.int32, device=DEV)
    nxt = torch.zeros(1, dtype=torch.int32, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    abi.mrope_positions(t, len(ids), grids, 300, pos, nxt, status)
    assert status.item() == 0
    assert np.array_equal(pos.cpu().numpy().astype(np.float32), ref[:, 0].numpy())
    abi.mrope_positions(t, 0, grids, 300, pos, nxt, status)
    assert nxt.item() == 0
    bad = torch.tensor([[1, 30, 32]] * 300, dtype=torch.int32, device=DEV)
    abi.mrope_positions(t, len(ids), bad, 300, pos, nxt, status)
    assert status.item() & 1


# ------------------------------------------------------------------------------------------------ compaction + RoPE
def _oracle_compact(arena_cpu, gather, pos_int, tables):
    """new K_raw / V = gather; K_rot = reference mRoPE of the gathered raw keys."""
    idx = torch.tensor(gather, dtype=torch.long)
    k_raw = arena_cpu[0, 0].index_select(2, idx)        # [L,H,n,D]
    v = arena_cpu[0, 2].index_select(2, idx)
    cos, sin = cos_sin_at(tables, pos_int)
    L, H, n, D = k_raw.shape
    _, k_rot = o_pos.apply_mrope(k_raw.reshape(1, L * H, n, D)[:, :1, :1], k_raw.reshape(1, L * H, n, D), cos, sin, SECTION)
    return k_raw, k_rot.reshape(L, H, n, D), v


@pytest.mark.parametrize("stream_idx,round_idx", [(0, -1), (0, 16), (1, -1), (2, -1), (3, -1), (4, -1)])
def test_kv_compact_rope_bit_exact(abi, tables, golden_dir, stream_idx, round_idx):
    streams = json.load(open(os.path.join(golden_dir, "eviction_streams.json")))
    rec = streams[stream_idx]["rounds"][round_idx]
    runs_rle = rec["cache_runs"]                      # [(src_start, len)] from the reference itself
    gather = [s + i for s, l in runs_rle for i in range(l)]
    n_src = rec["before_len"]
    cap = ((n_src + 127) // 128) * 128 + 128
    L, H = 2, 2
    arena_cpu, arena_dev, view = make_arena(abi, L, H, cap, seed=stream_idx)
    runs, d = [], 0
    for s, l in runs_rle:
        runs.append((s, d, l))
        d += l
    n = len(gather)
    # contiguous positions of the pruned sequence (any valid positions exercise the kernel; use real ones)
    g = torch.Generator().manual_seed(5)
    pos_i = torch.sort(torch.randint(0, 2000, (3, cap), generator=g), dim=1).values.to(torch.int32)
    pos_dev = pos_i.to(DEV)
    cos, sin = tables
    abi.kv_compact_rope(view, 0, runs, pos_dev, cos.to(DEV), sin.to(DEV), SECTION)
    torch.cuda.synchronize()
    got = arena_dev.cpu()
    k_raw, k_rot, v = _oracle_compact(arena_cpu, gather, pos_i[:, :n], tables)
    assert torch.equal(bits(got[1, 0, :, :, :n]), bits(k_raw))
    assert torch.equal(bits(got[1, 2, :, :, :n]), bits(v))
    assert torch.equal(bits(got[1, 1, :, :, :n]), bits(k_rot))
    assert torch.equal(bits(got[0]), bits(arena_cpu[0])), "source buffer must be untouched"
    assert (got[1, :, :, :, n:] == 0).all(), "rows past the retained length must not be written"


def test_kv_compact_rope_argument_errors(abi, tables):
    _, _, view = make_arena(abi, 1, 1, 256, fill=False)
    pos = torch.zeros(3, 256, dtype=torch.int32, device=DEV)
    cos, sin = (t.to(DEV) for t in tables)
    with pytest.raises(RuntimeError, match="contiguously"):
        abi.kv_compact_rope(view, 0, [(0, 5, 10)], pos, cos, sin, SECTION)
    with pytest.raises(RuntimeError, match="capacity"):
        abi.kv_compact_rope(view, 0, [(250, 0, 10)], pos, cos, sin, SECTION)
    with pytest.raises(RuntimeError, match="SVLM_MAX_RUNS"):
        abi.kv_compact_rope(view, 0, [(i, i, 1) for i in range(40)], pos, cos, sin, SECTION)
    abi.kv_compact_rope(view, 0, [], pos, cos, sin, SECTION)      # empty gather is a no-op


def test_qkv_rope_append_bit_exact(abi, tables):
    L, H, cap, Hq, T, start = 2, 2, 512, 4, 37, 200
    arena_cpu, arena_dev, view = make_arena(abi, L, H, cap, seed=9)
    arena_dev[1].copy_(arena_dev[0])
    g = torch.Generator().manual_seed(4)
    qkv = (torch.randn(T, (Hq + 2 * H) * 128, generator=g)).to(torch.bfloat16)
    qkv_d = qkv.to(DEV)
    q, k, v = qkv_d[:, :Hq * 128], qkv_d[:, Hq * 128:(Hq + H) * 128], qkv_d[:, (Hq + H) * 128:]
    pos_i = torch.sort(torch.randint(0, 3000, (3, cap), generator=g), dim=1).values.to(torch.int32)
    cos, sin = tables
    q_out = torch.zeros(T, Hq, 128, dtype=torch.bfloat16, device=DEV)
    abi.qkv_rope_append(view, 1, 1, q, k, v, T, Hq, q_out, start, pos_i.to(DEV), cos.to(DEV), sin.to(DEV), SECTION)
    torch.cuda.synchronize()
    c, s = cos_sin_at(tables, pos_i[:, start:start + T])
    q_ref, k_ref = o_pos.apply_mrope(qkv[:, :Hq * 128].view(1, T, Hq, 128).transpose(1, 2),
                                     qkv[:, Hq * 128:(Hq + H) * 128].view(1, T, H, 128).transpose(1, 2), c, s, SECTION)
    got = arena_dev.cpu()
    assert torch.equal(bits(q_out.cpu()), bits(q_ref[0].transpose(0, 1)))
    assert torch.equal(bits(got[1, 1, 1, :, start:start + T]), bits(k_ref[0]))
    assert torch.equal(bits(got[1, 0, 1, :, start:start + T]), bits(qkv[:, Hq * 128:(Hq + H) * 128].view(T, H, 128).transpose(0, 1)))
    assert torch.equal(bits(got[1, 2, 1, :, start:start + T]), bits(qkv[:, (Hq + H) * 128:].view(T, H, 128).transpose(0, 1)))
    # everything else untouched
    got[1, :, 1, :, start:start + T] = arena_cpu[0, :, 1, :, start:start + T]
    assert torch.equal(bits(got[1]), bits(arena_cpu[0]))
    # decode form: start and text position read from device memory
    start_dev = torch.tensor([start + T], dtype=torch.int32, device=DEV)
    tpos = torch.tensor([1234], dtype=torch.int32, device=DEV)
    q1 = torch.zeros(1, Hq, 128, dtype=torch.bfloat16, device=DEV)
    abi.qkv_rope_append(view, 1, 0, q[:1], k[:1], v[:1], 1, Hq, q1, 0, None, cos.to(DEV), sin.to(DEV), SECTION,
                        start_dev=start_dev, text_pos_dev=tpos)
    c, s = cos_sin_at(tables, torch.full((3, 1), 1234))
    q_ref, k_ref = o_pos.apply_mrope(qkv[:1, :Hq * 128].view(1, 1, Hq, 128).transpose(1, 2),
                                     qkv[:1, Hq * 128:(Hq + H) * 128].view(1, 1, H, 128).transpose(1, 2), c, s, SECTION)
    assert torch.equal(bits(q1.cpu()), bits(q_ref[0].transpose(0, 1)))
    assert torch.equal(bits(arena_dev[1, 1, 0, :, start + T].cpu()), bits(k_ref[0, :, 0]))


# ------------------------------------------------------------------------------------------------ attention
ATT_TOL = dict(rtol=2e-2, atol=6e-3)   # bf16 output, bf16 P; |out| <~ 1


def _attn_case(abi, Hq, H, Tq, S, splits, seed):
    cap = ((S + 127) // 128) * 128 + 128
    L = 2
    arena_cpu, arena_dev, view = make_arena(abi, L, H, cap, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    q = torch.randn(Tq, Hq, 128, generator=g).to(torch.bfloat16)
    out = torch.zeros(Tq, Hq, 128, dtype=torch.bfloat16, device=DEV)
    ws = None
    if splits > 1:
        ws = torch.zeros(abi.attn_workspace_bytes(Tq, Hq, 128, splits) // 4, dtype=torch.float32, device=DEV)
    layer = 1
    abi.attn_causal(view, 0, layer, q.to(DEV), out, Tq, Hq, S, 1.0 / math.sqrt(128), splits, ws)
    torch.cuda.synchronize()
    ref = o_attn.attention(q.transpose(0, 1), arena_cpu[0, 1, layer, :, :S], arena_cpu[0, 2, layer, :, :S], causal=True)
    return out.cpu().float().reshape(Tq, -1), ref.float()


@pytest.mark.parametrize("Hq,H,Tq,S,splits", [
    (4, 2, 1, 128, 1), (4, 2, 1, 300, 1), (4, 2, 64, 64, 1), (4, 2, 130, 700, 1),
    (28, 4, 276, 5619, 1), (28, 4, 276, 5619, 2), (28, 4, 1, 5343, 1), (28, 4, 1, 5343, 37), (28, 4, 1, 130, 8),
    (40, 8, 33, 1000, 3), (4, 2, 300, 300, 1),
])
def test_attn_causal(abi, Hq, H, Tq, S, splits):
    got, ref = _attn_case(abi, Hq, H, Tq, S, splits, seed=Hq + Tq + S)
    torch.testing.assert_close(got, ref, **ATT_TOL)


def test_attn_decode_kv_len_from_device(abi):
    Hq, H, S = 28, 4, 4000
    cap = 4224
    arena_cpu, arena_dev, view = make_arena(abi, 1, H, cap, seed=1)
    q = torch.randn(1, Hq, 128, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    out = torch.zeros(1, Hq, 128, dtype=torch.bfloat16, device=DEV)
    splits = 16
    ws = torch.zeros(abi.attn_workspace_bytes(1, Hq, 128, splits) // 4, dtype=torch.float32, device=DEV)
    kv = torch.tensor([S], dtype=torch.int32, device=DEV)
    abi.attn_causal(view, 0, 0, q.to(DEV), out, 1, Hq, cap, 1.0 / math.sqrt(128), splits, ws, kv_len_dev=kv)
    ref = o_attn.attention(q.transpose(0, 1), arena_cpu[0, 1, 0, :, :S], arena_cpu[0, 2, 0, :, :S], causal=True)
    torch.testing.assert_close(out.cpu().float().reshape(1, -1), ref.float(), **ATT_TOL)


@pytest.mark.parametrize("heads,D,n,cu", [
    (16, 80, 1024, list(range(0, 1025, 64))), (16, 80, 1024, [0, 1024]), (4, 80, 1024, list(range(0, 1025, 64))),
    (4, 80, 300, [0, 64, 100, 228, 300]), (2, 128, 256, [0, 256]),
])
def test_vit_rope_and_attention(abi, heads, D, n, cu):
    g = torch.Generator().manual_seed(n + heads)
    qkv = torch.randn(n, 3, heads, D, generator=g).to(torch.bfloat16)
    ang = torch.rand(n, D // 2, generator=g) * 6.0
    emb = torch.cat((ang, ang), dim=-1)
    cos, sin = emb.cos().contiguous(), emb.sin().contiguous()
    qkv_d = qkv.to(DEV)
    abi.vit_rope(qkv_d, cos.to(DEV), sin.to(DEV), n, heads, D)
    q_ref, k_ref = o_vit.apply_rope_vision(qkv[:, 0], qkv[:, 1], cos, sin)
    got = qkv_d.cpu()
    assert torch.equal(bits(got[:, 0]), bits(q_ref)) and torch.equal(bits(got[:, 1]), bits(k_ref))
    assert torch.equal(bits(got[:, 2]), bits(qkv[:, 2]))
    out = torch.zeros(n, heads, D, dtype=torch.bfloat16, device=DEV)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    abi.vit_attn(qkv_d, cu_t, out, n, heads, D, D ** -0.5)
    torch.cuda.synchronize()
    ref = o_attn.varlen_attention(q_ref, k_ref, qkv[:, 2], cu)
    torch.testing.assert_close(out.cpu().float(), ref.float(), **ATT_TOL)


# ------------------------------------------------------------------------------------------------ decode GEMV
@pytest.mark.parametrize("n,k", [(4608, 3584), (3584, 3584), (3584, 18944), (1536, 512), (512, 1024), (37, 64)])
def test_gemv_plain_and_residual(abi, n, k):
    g = torch.Generator().manual_seed(n + k)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(torch.bfloat16)
    x = torch.randn(k, generator=g).to(torch.bfloat16)
    b = (torch.randn(n, generator=g) * 0.1).to(torch.bfloat16)
    res = torch.randn(n, generator=g).to(torch.bfloat16)
    nw = (1 + 0.1 * torch.randn(k, generator=g)).to(torch.bfloat16)
    wd, xd = w.to(DEV), x.to(DEV)
    y = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    abi.gemv_bf16(wd, xd, y, 0, bias=b.to(DEV))
    ref = torch.nn.functional.linear(x.float(), w.float(), b.float()).to(torch.bfloat16)
    torch.testing.assert_close(y.cpu().float(), ref.float(), rtol=1.6e-2, atol=1e-2)
    # RMSNorm prologue + residual epilogue, in place on the residual
    yr = res.to(DEV).clone()
    abi.gemv_bf16(wd, xd, yr, 1, norm_weight=nw.to(DEV), eps=1e-6, residual=yr)
    xn = o_vit.rmsnorm(x[None], nw, 1e-6)[0]
    ref = (res.float() + torch.nn.functional.linear(xn.float(), w.float()).to(torch.bfloat16).float()).to(torch.bfloat16)
    torch.testing.assert_close(yr.cpu().float(), ref.float(), rtol=1.6e-2, atol=2e-2)


@pytest.mark.parametrize("n,k", [(18944, 3584), (1024, 512), (100, 64)])
def test_gemv_swiglu(abi, n, k):
    g = torch.Generator().manual_seed(n)
    w = (torch.randn(2 * n, k, generator=g) / math.sqrt(k)).to(torch.bfloat16)
    x = torch.randn(k, generator=g).to(torch.bfloat16)
    nw = (1 + 0.1 * torch.randn(k, generator=g)).to(torch.bfloat16)
    y = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    abi.gemv_bf16(w.to(DEV), x.to(DEV), y, 2, norm_weight=nw.to(DEV), eps=1e-6)
    xn = o_vit.rmsnorm(x[None], nw, 1e-6)[0]
    gu = torch.nn.functional.linear(xn, w)
    ref = torch.nn.functional.silu(gu[:n]) * gu[n:]
    torch.testing.assert_close(y.cpu().float(), ref.float(), rtol=2e-2, atol=1e-2)


def test_lm_head_argmax(abi):
    g = torch.Generator().manual_seed(0)
    n, k = 152064, 3584
    w = (torch.randn(n, k, generator=g) * 0.02).to(torch.bfloat16)
    nw = (1 + 0.1 * torch.randn(k, generator=g)).to(torch.bfloat16)
    sup = torch.zeros(n, dtype=torch.uint8)
    sup[151643:] = 1
    key = torch.zeros(1, dtype=torch.int64, device=DEV)
    tok = torch.zeros(1, dtype=torch.int64, device=DEV)
    cnt = torch.tensor([100, 0, 7], dtype=torch.int32, device=DEV)
    wd, nwd, supd = w.to(DEV), nw.to(DEV), sup.to(DEV)
    for trial in range(4):
        x = torch.randn(k, generator=g).to(torch.bfloat16)
        abi.lm_head_argmax(wd, x.to(DEV), nwd, 1e-6, supd, key, tok, cnt[0:1], cnt[1:2], cnt[2:3])
        xn = o_vit.rmsnorm(x[None], nw, 1e-6)[0]
        logits = torch.nn.functional.linear(xn.float(), w.float()).to(torch.bfloat16).float()
        logits[151643:] = -float("inf")
        t = int(tok.item())
        # the same bf16 logit value as the reference's maximum (ties and 1-ulp accumulation differences allowed)
        assert logits[t] >= logits.max() - 0.016 * abs(float(logits.max())), (t, int(logits.argmax()))
        assert t < 151643
    assert cnt.tolist() == [104, 105, 11] and key.item() == 0


def test_add_rmsnorm_and_swiglu(abi):
    g = torch.Generator().manual_seed(8)
    for rows, cols in [(278, 3584), (1024, 1280), (3, 512), (5, 8192)]:
        a = torch.randn(rows, cols, generator=g).to(torch.bfloat16)
        b = torch.randn(rows, cols, generator=g).to(torch.bfloat16)
        w = (1 + 0.1 * torch.randn(cols, generator=g)).to(torch.bfloat16)
        s_out = torch.empty(rows, cols, dtype=torch.bfloat16, device=DEV)
        n_out = torch.empty(rows, cols, dtype=torch.bfloat16, device=DEV)
        abi.add_rmsnorm(a.to(DEV), b.to(DEV), w.to(DEV), s_out, n_out, 1e-6)
        ref_s = a + b
        assert torch.equal(bits(s_out.cpu()), bits(ref_s))
        torch.testing.assert_close(n_out.cpu().float(), o_vit.rmsnorm(ref_s, w, 1e-6).float(), rtol=1.6e-2, atol=1e-3)
    for rows, m in [(278, 18944), (1024, 3456), (2, 64)]:
        gu = torch.randn(rows, 2 * m, generator=g).to(torch.bfloat16)
        out = torch.empty(rows, m, dtype=torch.bfloat16, device=DEV)
        abi.swiglu(gu.to(DEV), out)
        ref = torch.nn.functional.silu(gu[:, :m]) * gu[:, m:]
        torch.testing.assert_close(out.cpu().float(), ref.float(), rtol=1.6e-2, atol=1e-3)


def test_gemv_qkv_rope_append_matches_unfused(abi, tables):
    """Fused decode QKV projection + mRoPE + arena append == GEMV followed by svlm_qkv_rope_append."""
    L, H, cap, Hq, hidden = 2, 4, 512, 28, 3584
    g = torch.Generator().manual_seed(21)
    w = (torch.randn((Hq + 2 * H) * 128, hidden, generator=g) / math.sqrt(hidden)).to(torch.bfloat16).to(DEV)
    b = (torch.randn((Hq + 2 * H) * 128, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    x = torch.randn(hidden, generator=g).to(torch.bfloat16).to(DEV)
    nw = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16).to(DEV)
    cos, sin = (t.to(DEV) for t in tables)
    row = torch.tensor([77], dtype=torch.int32, device=DEV)
    pos = torch.tensor([1234], dtype=torch.int32, device=DEV)
    _, a1, v1 = make_arena(abi, L, H, cap, seed=3)
    _, a2, v2 = make_arena(abi, L, H, cap, seed=3)
    qkv = torch.zeros((Hq + 2 * H) * 128, dtype=torch.bfloat16, device=DEV)
    abi.gemv_bf16(w, x, qkv, 0, nw, 1e-6, b)
    q2 = qkv.view(1, -1)
    q_ref = torch.zeros(1, Hq, 128, dtype=torch.bfloat16, device=DEV)
    abi.qkv_rope_append(v1, 0, 1, q2[:, :Hq * 128], q2[:, Hq * 128:(Hq + H) * 128], q2[:, (Hq + H) * 128:], 1, Hq, q_ref, 0, None, cos, sin,
                        SECTION, start_dev=row, text_pos_dev=pos)
    q_fused = torch.zeros(1, Hq, 128, dtype=torch.bfloat16, device=DEV)
    abi.gemv_qkv_rope_append(v2, 0, 1, w, x, nw, 1e-6, b, Hq, q_fused, row, pos, cos, sin)
    torch.cuda.synchronize()
    # the two GEMVs split K across warps differently (fp32 summation order), so a dot product can land on the other side of
    # a bf16 rounding boundary: equal to one bf16 ulp, and everything the kernel does not touch is bit-identical
    torch.testing.assert_close(q_fused.cpu().float(), q_ref.cpu().float(), rtol=8e-3, atol=8e-3)
    torch.testing.assert_close(a2.cpu().float(), a1.cpu().float(), rtol=8e-3, atol=8e-3)
    untouched = torch.ones(512, dtype=torch.bool)
    untouched[77] = False
    assert torch.equal(bits(a2[:, :, :, :, untouched].cpu()), bits(a1[:, :, :, :, untouched].cpu()))
    assert torch.equal(bits(a2[:, :, 0].cpu()), bits(a1[:, :, 0].cpu())), "other layers untouched"


# ------------------------------------------------------------------------------------------------ several streams per launch
class _FakeCache:
    """The two attributes `_abi.make_slots` reads from a StreamingCache."""

    def __init__(self, arena, cur):
        self.arena, self.cur = arena, cur


def _stream_setup(abi, n_streams, L, H, cap, kv_lens, curs):
    arenas, caches = [], []
    for b in range(n_streams):
        cpu, dev, _ = make_arena(abi, L, H, cap, seed=50 + b)
        dev[1].copy_(dev[0])
        arenas.append((cpu, dev))
        caches.append(_FakeCache(dev, curs[b]))
    rows = [torch.tensor([n], dtype=torch.int32, device=DEV) for n in kv_lens]
    after = [torch.tensor([n + 1], dtype=torch.int32, device=DEV) for n in kv_lens]
    pos = [torch.tensor([1000 + 17 * b], dtype=torch.int32, device=DEV) for b in range(n_streams)]
    return arenas, caches, rows, after, pos


@pytest.mark.gpu
def test_qkv_rope_append_streams_matches_single_stream(abi, tables):
    """Token b lands in stream b's arena exactly as the single-stream append would put it; nothing else is touched."""
    B, L, H, cap, Hq = 3, 2, 2, 512, 4
    kv_lens, curs = [200, 37, 511], [1, 0, 1]
    arenas, caches, rows, after, pos = _stream_setup(abi, B, L, H, cap, kv_lens, curs)
    ref = [aligned_clone(a[1]) for a in arenas]
    g = torch.Generator().manual_seed(8)
    qkv = torch.randn(B, (Hq + 2 * H) * 128, generator=g).to(torch.bfloat16).to(DEV)
    q, k, v = qkv[:, :Hq * 128], qkv[:, Hq * 128:(Hq + H) * 128], qkv[:, (Hq + H) * 128:]
    cos, sin = (t.to(DEV) for t in tables)
    layer = 1
    q_ref = torch.zeros(B, Hq, 128, dtype=torch.bfloat16, device=DEV)
    for b in range(B):
        abi.qkv_rope_append(abi.ArenaView(ref[b]), curs[b], layer, q[b:b + 1], k[b:b + 1], v[b:b + 1], 1, Hq, q_ref[b:b + 1], 0, None,
                            cos, sin, SECTION, start_dev=rows[b], text_pos_dev=pos[b])
    q_out = torch.zeros(B, Hq, 128, dtype=torch.bfloat16, device=DEV)
    slots = abi.make_slots(caches, rows, pos, after)
    abi.qkv_rope_append_streams(abi.ArenaView(arenas[0][1]), slots, layer, q, k, v, Hq, q_out, cos, sin, SECTION)
    torch.cuda.synchronize()
    assert torch.equal(bits(q_out.cpu()), bits(q_ref.cpu()))
    for b in range(B):
        assert torch.equal(bits(arenas[b][1].cpu()), bits(ref[b].cpu())), f"stream {b}"
        assert not torch.equal(bits(arenas[b][1].cpu()), bits(arenas[b][0]))       # the row really was written


@pytest.mark.gpu
@pytest.mark.parametrize("B,kv_lens,splits", [(3, [300, 5343, 129], 8), (8, [5789, 5700, 4100, 128, 1, 6000, 5789, 2500], 16),
                                              (2, [1000, 77], 1),
                                              (16, [5789, 12, 4100, 128, 1, 6000, 5789, 2500, 300, 5343, 129, 3000, 2047, 2048, 2049, 640], 2),
                                              (12, [700 + 411 * b for b in range(12)], 4)])
def test_attn_decode_streams_matches_single_stream(abi, B, kv_lens, splits):
    """One launch over B streams == B single-stream launches, bit for bit (same kernel, same split boundaries per stream)."""
    Hq, H, L = 28, 4, 2
    cap = 6144
    curs = [b & 1 for b in range(B)]
    arenas, caches, rows, after, pos = _stream_setup(abi, B, L, H, cap, kv_lens, curs)
    kv = [torch.tensor([n], dtype=torch.int32, device=DEV) for n in kv_lens]
    q = torch.randn(B, Hq, 128, generator=torch.Generator().manual_seed(B)).to(torch.bfloat16).to(DEV)
    layer, scale = 1, 1.0 / math.sqrt(128)
    ws1 = torch.zeros(abi.attn_workspace_bytes(1, Hq, 128, splits) // 4 + 16, dtype=torch.float32, device=DEV) if splits > 1 else None
    ref = torch.zeros(B, Hq, 128, dtype=torch.bfloat16, device=DEV)
    for b in range(B):
        abi.attn_causal(abi.ArenaView(arenas[b][1]), curs[b], layer, q[b:b + 1], ref[b:b + 1], 1, Hq, cap, scale, splits, ws1, kv_len_dev=kv[b])
    wsB = torch.zeros(abi.attn_workspace_bytes(B, Hq, 128, splits) // 4 + 16, dtype=torch.float32, device=DEV) if splits > 1 else None
    out = torch.zeros(B, Hq, 128, dtype=torch.bfloat16, device=DEV)
    slots = abi.make_slots(caches, rows, pos, kv)
    abi.attn_decode_streams(abi.ArenaView(arenas[0][1]), slots, layer, q, out, Hq, cap, scale, splits, wsB)
    torch.cuda.synchronize()
    if splits > 1 and splits <= 8 and splits & (splits - 1) == 0:
        # the batched launch merges these split counts inside thread-block clusters, the single-stream launch through the merge
        # kernel: same partial results, a different order of the fp32 merge arithmetic (<= 1 bf16 ulp of a few outputs)
        torch.testing.assert_close(out.cpu().float(), ref.cpu().float(), rtol=8e-3, atol=2e-3)
    else:
        assert torch.equal(bits(out.cpu()), bits(ref.cpu()))
    # and against the definition for one of the streams
    b = 0
    cpu = arenas[b][0]
    want = o_attn.attention(q[b:b + 1].cpu().transpose(0, 1), cpu[0, 1, layer, :, :kv_lens[b]], cpu[0, 2, layer, :, :kv_lens[b]], causal=True)
    torch.testing.assert_close(out[b].cpu().float().reshape(1, -1), want.float(), **ATT_TOL)


@pytest.mark.gpu
def test_streams_argument_errors(abi):
    cpu, dev, view = make_arena(abi, 1, 2, 256, seed=3)
    one = torch.zeros(1, dtype=torch.int32, device=DEV)
    too_many = abi.MAX_STREAMS + 1
    q = torch.zeros(too_many, 4, 128, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(ValueError):
        abi.make_slots([_FakeCache(dev, 0)] * too_many, [one] * too_many, [one] * too_many, [one] * too_many)
    slots = abi.make_slots([_FakeCache(dev, 0)], [one], [one], [one])
    with pytest.raises(RuntimeError, match="workspace"):
        abi.attn_decode_streams(view, slots, 0, q[:1], q[:1].clone(), 4, 256, 0.1, 4, None)
    with pytest.raises(RuntimeError, match="layer"):
        abi.attn_decode_streams(view, slots, 5, q[:1], q[:1].clone(), 4, 256, 0.1, 1, None)


# ------------------------------------------------------------------------------------------------ frame ingest (row f3)
@pytest.mark.gpu
def test_frames_to_patches_bit_exact(abi, golden_dir):
    from oracle import frames as o_frames
    from streaming_vlm_b200 import synthetic
    z = np.load(os.path.join(golden_dir, "frame_ingest.npz"))
    cases = [torch.from_numpy(z[n + "__frames_thwc"]) for n in ("small", "wide")]
    cases.append(synthetic.frames(5, 448, 448).permute(0, 2, 3, 1).contiguous())
    cases.append(synthetic.frames(6, 56, 56, n=6).permute(0, 2, 3, 1).contiguous())            # three temporal patches
    for fr in cases:
        want, grid = o_frames.frames_to_patches(fr)
        got32, g32 = abi.frames_to_patches(fr.to(DEV), dtype=torch.float32)
        got16, g16 = abi.frames_to_patches(fr.to(DEV), dtype=torch.bfloat16)
        torch.cuda.synchronize()
        assert g32 == grid and g16 == grid
        assert torch.equal(got32.cpu().view(torch.int32), want.view(torch.int32))
        assert torch.equal(bits(got16.cpu()), bits(want.to(torch.bfloat16)))
    # and against the HF processor's own output
    got, _ = abi.frames_to_patches(cases[0].to(DEV), dtype=torch.float32)
    assert np.array_equal(got.cpu().numpy().view(np.int32), z["small__pixel_values"].view(np.int32))
    with pytest.raises(RuntimeError, match="multiple of patch"):
        abi.frames_to_patches(torch.zeros(2, 30, 56, 3, dtype=torch.uint8, device=DEV))
    with pytest.raises(RuntimeError, match="temporal_patch"):
        abi.frames_to_patches(torch.zeros(3, 56, 56, 3, dtype=torch.uint8, device=DEV))


# ------------------------------------------------------------------------------------------------ against the reference's own attention library
def _flash_attn():
    try:
        from flash_attn import flash_attn_func, flash_attn_varlen_func
        return flash_attn_func, flash_attn_varlen_func
    except Exception:        # noqa: BLE001  (library not in the image)
        return None, None


@pytest.mark.gpu
@pytest.mark.parametrize("Tq,S,splits", [(276, 5619, 2), (300, 5789, 1), (1, 5343, 16), (1, 40000, 37), (130, 700, 1)])
def test_attn_causal_matches_flash_attn(abi, Tq, S, splits):
    """The reference's attention path, run as the reference runs it (`repeat_kv` x7, `flash_attn_func(causal=True)`: bottom-right
    aligned, scale 1/sqrt(128); language_forward.py:199-252) on the same K_rot / V: outputs agree within the bf16 tolerance."""
    flash_attn_func, _ = _flash_attn()
    if flash_attn_func is None:
        pytest.skip("flash_attn is not installed")
    Hq, H, L = 28, 4, 1
    cap = ((S + 127) // 128) * 128 + 128
    _, arena_dev, view = make_arena(abi, L, H, cap, seed=S)
    q = torch.randn(Tq, Hq, 128, generator=torch.Generator().manual_seed(Tq)).to(torch.bfloat16).to(DEV)
    out = torch.zeros_like(q)
    ws = torch.zeros(abi.attn_workspace_bytes(Tq, Hq, 128, splits) // 4 + 16, dtype=torch.float32, device=DEV) if splits > 1 else None
    abi.attn_causal(view, 0, 0, q, out, Tq, Hq, S, 1.0 / math.sqrt(128), splits, ws)
    k = arena_dev[0, 1, 0, :, :S].transpose(0, 1).repeat_interleave(Hq // H, dim=1)[None].contiguous()      # [1, S, Hq, D]
    v = arena_dev[0, 2, 0, :, :S].transpose(0, 1).repeat_interleave(Hq // H, dim=1)[None].contiguous()
    ref = flash_attn_func(q[None], k, v, causal=True)[0]
    torch.testing.assert_close(out.float().cpu(), ref.float().cpu(), **ATT_TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("cu", [list(range(0, 1025, 64)), [0, 1024], [0, 64, 100, 228, 300]])
def test_vit_attn_matches_flash_attn_varlen(abi, cu):
    """ViT attention against `flash_attn_varlen_func` (vision_forward.py:32-35) on the same rotated q, k and v."""
    _, flash_attn_varlen_func = _flash_attn()
    if flash_attn_varlen_func is None:
        pytest.skip("flash_attn is not installed")
    heads, D, n = 16, 80, cu[-1]
    qkv = torch.randn(n, 3, heads, D, generator=torch.Generator().manual_seed(n)).to(torch.bfloat16).to(DEV)
    out = torch.zeros(n, heads, D, dtype=torch.bfloat16, device=DEV)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=DEV)
    abi.vit_attn(qkv, cu_t, out, n, heads, D, D ** -0.5)
    max_len = max(b - a for a, b in zip(cu[:-1], cu[1:]))
    ref = flash_attn_varlen_func(qkv[:, 0].contiguous(), qkv[:, 1].contiguous(), qkv[:, 2].contiguous(), cu_t, cu_t, max_len, max_len)
    torch.testing.assert_close(out.float().cpu(), ref.float().cpu(), **ATT_TOL)


@pytest.mark.gpu
def test_frames_resize_bicubic(abi, golden_dir):
    """Kernel == oracle bit for bit (same tap tables, same fp32 accumulation order); == torchvision within 1 LSB."""
    from oracle import frames as o_frames
    from streaming_vlm_b200 import synthetic
    z = np.load(os.path.join(golden_dir, "frame_ingest.npz"))
    for name in ("down", "mixed", "up"):
        clip, want = torch.from_numpy(z[f"resize_{name}__in_thwc"]), torch.from_numpy(z[f"resize_{name}__out_thwc"])
        got = abi.frames_resize_bicubic(clip.to(DEV), want.shape[1], want.shape[2]).cpu()
        assert torch.equal(got, o_frames.resize_bicubic_aa(clip, want.shape[1], want.shape[2])), name
        d = (got.int() - want.int()).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3
    # a 720p clip to the size smart_resize picks, then straight into the patchifier
    clip = synthetic.frames(9, 360, 640).permute(0, 2, 3, 1).contiguous()
    from streaming_vlm_b200.utils.frames import smart_resize
    h2, w2 = smart_resize(360, 640, 28, 128 * 28 * 28, 256 * 28 * 28)
    small = abi.frames_resize_bicubic(clip.to(DEV), h2, w2)
    assert torch.equal(small.cpu(), o_frames.resize_bicubic_aa(clip, h2, w2))
    rows, grid = abi.frames_to_patches(small)
    want_rows, want_grid = o_frames.frames_to_patches(small.cpu())
    assert grid == want_grid and torch.equal(bits(rows.cpu()), bits(want_rows.to(torch.bfloat16)))


@pytest.mark.gpu
def test_torch_custom_ops_layer(abi):
    """The thin `torch.ops.svlm.*` layer (CUDA key only) returns what the ctypes wrappers return."""
    abi.register_torch_ops()
    g = torch.Generator().manual_seed(0)
    a = torch.randn(5, 512, generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randn(5, 512, generator=g).to(torch.bfloat16).to(DEV)
    wn = torch.rand(512, generator=g).to(torch.bfloat16).to(DEV)
    s_, h = torch.ops.svlm.add_rmsnorm(a, b, wn, 1e-6)
    s2, h2 = torch.empty_like(a), torch.empty_like(a)
    abi.add_rmsnorm(a, b, wn, s2, h2, 1e-6)
    assert torch.equal(bits(s_), bits(s2)) and torch.equal(bits(h), bits(h2))
    gu = torch.randn(5, 1024, generator=g).to(torch.bfloat16).to(DEV)
    out = torch.empty(5, 512, dtype=torch.bfloat16, device=DEV)
    abi.swiglu(gu, out)
    assert torch.equal(bits(torch.ops.svlm.swiglu(gu)), bits(out))
    w = (torch.randn(256, 512, generator=g) * 0.05).to(torch.bfloat16).to(DEV)
    x = torch.randn(512, generator=g).to(torch.bfloat16).to(DEV)
    y = torch.empty(256, dtype=torch.bfloat16, device=DEV)
    abi.gemv_bf16(w, x, y, 0, wn, 1e-6)
    assert torch.equal(bits(torch.ops.svlm.gemv(w, x, 0, wn, 1e-6, None, None)), bits(y))
    assert torch.ops.svlm.gemv(w, x, 2, wn, 1e-6, None, None).shape == (128,)
    fr = torch.randint(0, 256, (2, 56, 84, 3), dtype=torch.uint8, generator=g).to(DEV)
    assert torch.equal(bits(torch.ops.svlm.frames_to_patches(fr, 14, 2, 2)), bits(abi.frames_to_patches(fr)[0]))
    assert torch.equal(torch.ops.svlm.frames_resize_bicubic(fr, 28, 56), abi.frames_resize_bicubic(fr, 28, 56))
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.svlm.swiglu(gu.cpu())              # no CPU kernel is registered: there is no fallback


# ------------------------------------------------------------------------------------------------ embed + video splice
def test_embed_tokens_equals_embedding_plus_masked_scatter(abi):
    """svlm_embed_tokens == F.embedding followed by the reference's masked_scatter of the video embeddings
    (qwen2_5/model_forward.py:58, 87-93), for scattered and contiguous placeholder layouts; the status word flags a placeholder
    count above the number of feature rows and ids outside the vocabulary."""
    g = torch.Generator().manual_seed(11)
    vocab, hidden, vid = 5000, 512, 4999
    emb = torch.randn(vocab, hidden, generator=g).to(torch.bfloat16)
    for layout in ("block", "scattered", "none"):
        ids = torch.randint(0, vocab - 1, (301,), generator=g)
        if layout == "block":
            ids[20:276] = vid
        elif layout == "scattered":
            ids[torch.randperm(301, generator=g)[:77]] = vid
        n_vid = int((ids == vid).sum())
        video = torch.randn(max(n_vid, 1), hidden, generator=g).to(torch.bfloat16)
        want = torch.nn.functional.embedding(ids, emb)
        if n_vid:
            mask = (ids == vid)[:, None].expand_as(want)
            want = want.masked_scatter(mask, video[:n_vid])
        out = torch.empty(301, hidden, dtype=torch.bfloat16, device=DEV)
        status = torch.zeros(1, dtype=torch.int32, device=DEV)
        abi.embed_tokens(ids.to(DEV), 301, emb.to(DEV), video.to(DEV) if n_vid else None, n_vid, out, status, vid)
        assert torch.equal(bits(out.cpu()), bits(want)), layout
        assert int(status) == 0
    # more placeholders than feature rows -> bit 8; an id beyond the vocabulary -> bit 16
    ids = torch.tensor([1, vid, vid, vid, 2], dtype=torch.long)
    out = torch.empty(5, hidden, dtype=torch.bfloat16, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    abi.embed_tokens(ids.to(DEV), 5, emb.to(DEV), torch.zeros(2, hidden, dtype=torch.bfloat16, device=DEV), 2, out, status, vid)
    assert int(status) & 8
    status.zero_()
    abi.embed_tokens(torch.tensor([1, vocab + 5], device=DEV), 2, emb.to(DEV), None, 0, out, status, vid)
    assert int(status) & 16


def test_qkv_rope_append_without_queries(abi, tables):
    """q_heads = 0: the append-only form `StreamingCache.update` uses -- K_raw / K_rot / V rows, bit-exact vs the oracle's mRoPE."""
    L, H, cap, t, start = 2, 2, 256, 37, 50
    _, dev, view = make_arena(abi, L, H, cap, fill=False)
    g = torch.Generator().manual_seed(3)
    k = (torch.randn(t, H * 128, generator=g) * 0.5).to(torch.bfloat16)
    v = (torch.randn(t, H * 128, generator=g) * 0.5).to(torch.bfloat16)
    pos = torch.zeros(3, cap, dtype=torch.int32)
    pos[:, start:start + t] = torch.randint(0, 4000, (3, t), generator=g, dtype=torch.int32)
    tc, ts = tables
    abi.qkv_rope_append(view, 0, 1, None, k.to(DEV), v.to(DEV), t, 0, None, start, pos.to(DEV), tc.to(DEV), ts.to(DEV), SECTION)
    got = dev.cpu()
    kk = k.view(t, H, 128).transpose(0, 1)
    cos, sin = cos_sin_at(tables, pos[:, start:start + t])
    _, want_rot = o_pos.apply_mrope(kk[None, :1, :1], kk[None], cos, sin, SECTION)
    assert torch.equal(bits(got[0, 0, 1, :, start:start + t]), bits(kk))
    assert torch.equal(bits(got[0, 1, 1, :, start:start + t]), bits(want_rot[0]))
    assert torch.equal(bits(got[0, 2, 1, :, start:start + t]), bits(v.view(t, H, 128).transpose(0, 1)))
    assert got[0, :, 0].abs().sum() == 0 and got[0, :, 1, :, :start].abs().sum() == 0          # nothing else was touched


# ------------------------------------------------------------------------------------------------ persistent GEMM epilogues
@pytest.mark.parametrize("m,n,k,tile", [(300, 1280, 1280, 0), (1024, 3840, 1280, 256), (77, 3584, 1176, 64), (256, 512, 5120, 128),
                                         (1024, 1280, 3456, 0), (640, 2560, 192, 256),
                                         # 1000 + BN: the CTA-pair kernel (cta_group::2, 256 x BN tiles), incl. ragged M / N tails
                                         (1024, 3840, 1280, 1256), (256, 5120, 5120, 1128), (300, 1288, 1280, 1064), (1000, 2600, 704, 1256),
                                         (2048, 6912, 1280, 1128)])
def test_gemm_ex_bias_and_residual(abi, m, n, k, tile):
    """Persistent tcgen05 GEMM (several tiles per CTA, two TMEM accumulators) vs fp32 reference of the same op order: plain, bias,
    residual add (`hidden = residual + bf16(linear)`), every tile width, ragged M / N / K tails."""
    g = torch.Generator().manual_seed(m + n)
    a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    bias = (torch.randn(n, generator=g) * 0.2).to(torch.bfloat16)
    res = torch.randn(m, n, generator=g).to(torch.bfloat16)
    lin = (a.float() @ w.float().t() + bias.float()).to(torch.bfloat16)
    out = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
    abi.gemm_ex(a.to(DEV), w.to(DEV), out, bias=bias.to(DEV), tile_n=tile)
    torch.testing.assert_close(out.cpu().float(), lin.float(), rtol=1.6e-2, atol=1e-2)
    abi.gemm_ex(a.to(DEV), w.to(DEV), out, bias=bias.to(DEV), residual=res.to(DEV), epilogue=abi.EPI_RESIDUAL, tile_n=tile)
    want = (lin.float() + res.float()).to(torch.bfloat16)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=1.6e-2, atol=1.6e-2)
    # in place on the residual stream
    x = res.to(DEV).clone()
    abi.gemm_ex(a.to(DEV), w.to(DEV), x, bias=bias.to(DEV), residual=x, epilogue=abi.EPI_RESIDUAL, tile_n=tile)
    assert torch.equal(bits(x.cpu()), bits(out.cpu()))


@pytest.mark.parametrize("m,width,k", [(300, 1024, 512), (1024, 3456, 1280), (17, 18944, 3584)])
def test_gemm_ex_swiglu(abi, m, width, k):
    """SwiGLU epilogue on gate / up rows interleaved in blocks of 64: == silu(bf16(x Wg^T)) * bf16(x Wu^T) with the reference's roundings."""
    g = torch.Generator().manual_seed(width)
    a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16)
    wg = (torch.randn(width, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    wu = (torch.randn(width, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    gate = (a.float() @ wg.float().t()).to(torch.bfloat16)
    up = (a.float() @ wu.float().t()).to(torch.bfloat16)
    want = (torch.nn.functional.silu(gate.float()).to(torch.bfloat16).float() * up.float()).to(torch.bfloat16)
    w = abi.interleave_gate_up(wg, wu).to(DEV)
    out = torch.empty(m, width, dtype=torch.bfloat16, device=DEV)
    abi.gemm_ex(a.to(DEV), w, out, epilogue=abi.EPI_SWIGLU)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=2e-2, atol=1e-2)


# ---------------------------------------------------------------------------------------------------------------------------------
# svlm_gemm_tokens: few tokens x whole weight matrix (swap-AB tcgen05, split-K over a cluster), the batched decode step's projections
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("t,n,k,splits", [
    (8, 4608, 3584, 0), (8, 3584, 3584, 0), (8, 3584, 18944, 0),          # 7B q/k/v, o_proj, down_proj: auto split (4 / 5 / 5)
    (1, 1024, 512, 0), (3, 1024, 512, 1), (16, 512, 1024, 8),             # tiny geometry, explicit split counts
    (17, 200, 520, 3), (32, 384, 4096, 7), (40, 136, 1096, 2), (64, 1280, 640, 0),   # ragged N / K tails, 32- and 64-column accumulators
])
def test_gemm_tokens_bias_and_residual(abi, t, n, k, splits):
    g = torch.Generator().manual_seed(t + n + k)
    x = (torch.randn(t, k, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    bias = (torch.randn(n, generator=g) * 0.2).to(torch.bfloat16)
    res = torch.randn(t, n, generator=g).to(torch.bfloat16)
    lin = (x.float() @ w.float().t() + bias.float()).to(torch.bfloat16)
    xd, wd = x.to(DEV), w.to(DEV)
    out = torch.full((t, n), 7.0, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(xd, wd, out, bias=bias.to(DEV), splits=splits)
    torch.testing.assert_close(out.cpu().float(), lin.float(), rtol=1.6e-2, atol=1e-2)
    nob = torch.empty_like(out)
    abi.gemm_tokens(xd, wd, nob, splits=splits)
    torch.testing.assert_close(nob.cpu().float(), (x.float() @ w.float().t()).to(torch.bfloat16).float(), rtol=1.6e-2, atol=1e-2)
    abi.gemm_tokens(xd, wd, out, bias=bias.to(DEV), residual=res.to(DEV), epilogue=abi.EPI_RESIDUAL, splits=splits)
    want = (lin.float() + res.float()).to(torch.bfloat16)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=1.6e-2, atol=1.6e-2)
    # in place on the residual stream, and deterministic (the split-K partial tiles are summed in rank order)
    r2 = res.to(DEV).clone()
    abi.gemm_tokens(xd, wd, r2, bias=bias.to(DEV), residual=r2, epilogue=abi.EPI_RESIDUAL, splits=splits)
    assert torch.equal(bits(r2.cpu()), bits(out.cpu()))


def test_gemm_tokens_split_counts_agree(abi):
    """Every split count gives the same result up to the fp32 summation order (<= 1 bf16 ulp), a row of tokens equals the GEMV."""
    g = torch.Generator().manual_seed(5)
    t, n, k = 8, 3584, 3584
    x = (torch.randn(t, k, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(torch.bfloat16).to(DEV)
    outs = []
    for s in range(1, 9):
        o = torch.empty(t, n, dtype=torch.bfloat16, device=DEV)
        abi.gemm_tokens(x, w, o, splits=s)
        outs.append(o.cpu().float())
    for o in outs[1:]:
        torch.testing.assert_close(o, outs[0], rtol=8e-3, atol=8e-3)
    y = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    abi.gemv_bf16(w, x[3].contiguous(), y, 0)
    torch.testing.assert_close(y.cpu().float(), outs[0][3], rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("t,width,k,splits", [(8, 18944, 3584, 0), (5, 1024, 512, 0), (16, 100, 200, 2), (33, 640, 320, 1)])
def test_gemm_tokens_swiglu(abi, t, width, k, splits):
    """SwiGLU epilogue on W = [gate rows ; up rows] (the engine's own layout): a tile pairs 64 gate rows with the 64 up rows of the same
    columns; == silu(bf16(x Wg^T)) * bf16(x Wu^T) with the reference's roundings."""
    g = torch.Generator().manual_seed(width + t)
    x = (torch.randn(t, k, generator=g) * 0.5).to(torch.bfloat16)
    wg = (torch.randn(width, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    wu = (torch.randn(width, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    gate = (x.float() @ wg.float().t()).to(torch.bfloat16)
    up = (x.float() @ wu.float().t()).to(torch.bfloat16)
    want = (torch.nn.functional.silu(gate.float()).to(torch.bfloat16).float() * up.float()).to(torch.bfloat16)
    out = torch.full((t, width), 7.0, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(x.to(DEV), torch.cat([wg, wu]).to(DEV), out, epilogue=abi.EPI_SWIGLU, splits=splits)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=2e-2, atol=1e-2)


def test_gemm_tokens_lm_head_rows(abi):
    """lm_head width (1188 row tiles, two CTAs per SM, four waves): argmax over the bf16 logits == the fp32 reference's."""
    g = torch.Generator().manual_seed(9)
    t, n, k = 8, 152064, 3584
    x = torch.randn(t, k, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(n, k, generator=g) * 0.02).to(torch.bfloat16).to(DEV)
    out = torch.empty(t, n, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(x, w, out)
    ref = (x.float() @ w.float().t()).to(torch.bfloat16)
    torch.testing.assert_close(out.float(), ref.float(), rtol=1.6e-2, atol=1e-2)


def test_gemm_tokens_rejects_bad_arguments(abi):
    x = torch.zeros(65, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(128, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="tokens"):
        abi.gemm_tokens(x, w, torch.zeros(65, 128, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(RuntimeError, match="residual"):
        abi.gemm_tokens(x[:4].contiguous(), w, torch.zeros(4, 128, dtype=torch.bfloat16, device=DEV), epilogue=abi.EPI_RESIDUAL)


@pytest.mark.parametrize("t,n,k,epi,splits", [(8, 4608, 3584, 0, 0), (8, 37888, 3584, 3, 0), (5, 1024, 512, 0, 2), (16, 2048, 520, 3, 1),
                                              (1, 152064, 3584, 0, 0), (33, 256, 1096, 0, 3)])
def test_gemm_tokens_fused_rmsnorm_equals_norm_kernel_then_gemm(abi, t, n, k, epi, splits):
    """RMSNorm fused into the GEMM (row statistics on the epilogue warps, X panels normalised in shared memory) is BIT-identical to
    svlm_add_rmsnorm followed by the plain GEMM: same summation order of the statistics, same rounding points."""
    g = torch.Generator().manual_seed(t * n + k)
    x = (torch.randn(t, k, generator=g) * 2.0).to(torch.bfloat16).to(DEV)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(torch.bfloat16).to(DEV)
    nw = (1 + 0.1 * torch.randn(k, generator=g)).to(torch.bfloat16).to(DEV)
    h = torch.empty_like(x)
    abi.add_rmsnorm(x, None, nw, None, h, 1e-6)
    ref_h = o_vit.rmsnorm(x.cpu(), nw.cpu(), 1e-6)
    torch.testing.assert_close(h.cpu().float(), ref_h.float(), rtol=8e-3, atol=8e-3)
    cols = n // 2 if epi == abi.EPI_SWIGLU else n
    want = torch.empty(t, cols, dtype=torch.bfloat16, device=DEV)
    got = torch.full((t, cols), 3.0, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(h, w, want, epilogue=epi, splits=splits)
    abi.gemm_tokens(x, w, got, epilogue=epi, splits=splits, norm_weight=nw, eps=1e-6)
    assert torch.equal(bits(got.cpu()), bits(want.cpu()))


@pytest.mark.parametrize("n_streams,hq,hkv,hidden", [(8, 28, 4, 3584), (3, 4, 2, 512), (1, 4, 2, 512), (16, 4, 2, 512), (13, 28, 4, 3584)])
def test_gemm_tokens_qkv_append_equals_unfused_chain(abi, tables, n_streams, hq, hkv, hidden):
    """RMSNorm + q/k/v projection + bias + mRoPE + append into each stream's arena in ONE launch == svlm_add_rmsnorm + GEMM +
    svlm_qkv_rope_append_streams, bit for bit (q_rot and every arena, including the rows nobody touches)."""
    L, cap = 2, 256
    g = torch.Generator().manual_seed(hidden + n_streams)
    n = (hq + 2 * hkv) * 128
    w = (torch.randn(n, hidden, generator=g) * hidden ** -0.5).to(torch.bfloat16).to(DEV)
    b = (torch.randn(n, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    nw = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16).to(DEV)
    x = torch.randn(n_streams, hidden, generator=g).to(torch.bfloat16).to(DEV)
    cos, sin = (t.to(DEV) for t in tables)
    rows = torch.tensor([5 + 13 * s for s in range(n_streams)], dtype=torch.int32, device=DEV)
    poss = torch.tensor([900 + 101 * s for s in range(n_streams)], dtype=torch.int32, device=DEV)

    class C:        # what make_slots reads of a cache
        pass

    def fresh():
        caches, arenas = [], []
        for s in range(n_streams):
            _, a, v = make_arena(abi, L, hkv, cap, seed=40 + s)
            c = C(); c.arena, c.cur, c.view = a, s % 2, v
            caches.append(c); arenas.append(a)
        slots = abi.make_slots(caches, list(rows.split(1)), list(poss.split(1)), list(rows.split(1)))
        return caches, arenas, slots

    c1, a1, s1 = fresh()
    c2, a2, s2 = fresh()
    h = torch.empty_like(x)
    abi.add_rmsnorm(x, None, nw, None, h, 1e-6)
    qkv = torch.empty(n_streams, n, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(h, w, qkv, bias=b)
    q_ref = torch.zeros(n_streams, hq, 128, dtype=torch.bfloat16, device=DEV)
    abi.qkv_rope_append_streams(c1[0].view, s1, 1, qkv[:, :hq * 128], qkv[:, hq * 128:(hq + hkv) * 128], qkv[:, (hq + hkv) * 128:], hq, q_ref,
                                cos, sin, SECTION)
    q_got = torch.zeros_like(q_ref)
    abi.gemm_tokens_qkv_append(c2[0].view, s2, 1, x, w, b, nw, 1e-6, hq, q_got, cos, sin)
    torch.cuda.synchronize()
    assert torch.equal(bits(q_got.cpu()), bits(q_ref.cpu()))
    for s in range(n_streams):
        assert torch.equal(bits(a2[s].cpu()), bits(a1[s].cpu())), f"arena of stream {s} differs"
    # and the unfused chain itself against the fp32 definition of the rotation for one stream
    assert not torch.equal(bits(a1[0][c1[0].cur, 1, 1, :, 5].cpu()), bits(a1[0][c1[0].cur, 0, 1, :, 5].cpu()))


@pytest.mark.parametrize("t,hidden,mlp", [(8, 3584, 18944), (3, 512, 1024)])
def test_gemm_tokens_statistics_hand_off(abi, t, hidden, mlp):
    """A residual-epilogue launch leaves the per-tile sums of squares of the rows it wrote; the next launch's fused RMSNorm takes its
    row statistics from them (one short read instead of a pass over X per CTA).  The table equals the fp32 definition; the consumer
    equals svlm_add_rmsnorm + GEMM up to the summation order of the statistic (<= 1 bf16 ulp of a few outputs)."""
    g = torch.Generator().manual_seed(hidden + t)
    a = (torch.randn(t, mlp, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    wd = (torch.randn(hidden, mlp, generator=g) * mlp ** -0.5).to(torch.bfloat16).to(DEV)
    x = (torch.randn(t, hidden, generator=g) * 3.0).to(torch.bfloat16).to(DEV)
    nw = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16).to(DEV)
    wgu = (torch.randn(2 * mlp, hidden, generator=g) * hidden ** -0.5).to(torch.bfloat16).to(DEV)
    tiles = -(-hidden // 128)
    stats = torch.full((tiles, t), -1.0, dtype=torch.float32, device=DEV)
    abi.gemm_tokens(a, wd, x, residual=x, epilogue=abi.EPI_RESIDUAL, stats_out=stats)
    want_stats = x.float().pow(2).view(t, tiles, 128).sum(-1).t()
    torch.testing.assert_close(stats, want_stats, rtol=1e-5, atol=1e-6)
    h = torch.empty_like(x)
    abi.add_rmsnorm(x, None, nw, None, h, 1e-6)
    want = torch.empty(t, mlp, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(h, wgu, want, epilogue=abi.EPI_SWIGLU)
    got = torch.empty_like(want)
    abi.gemm_tokens(x, wgu, got, epilogue=abi.EPI_SWIGLU, norm_weight=nw, eps=1e-6, stats_in=stats)
    torch.testing.assert_close(got.float(), want.float(), rtol=1.6e-2, atol=1e-2)
    frac = (bits(got.cpu()) != bits(want.cpu())).float().mean().item()
    assert frac < 0.02, f"{frac:.4f} of the outputs differ from the unfused chain"


@pytest.mark.parametrize("kv_after,splits", [(5789, 23), (129, 1), (128, 1), (257, 2), (256, 2), (1, 1), (4097, 16), (6000, 8)])
def test_attn_decode_early_kv_prefetch_is_bit_identical(abi, tables, kv_after, splits):
    """SVLM_ATTN_EARLY_KV: the decode attention reads the device-side length and requests the K/V tiles below the row being appended
    BEFORE its grid dependency resolves.  In the real chain (q/k/v GEMV appending row kv_len - 1 right before it, programmatic
    dependent launch, CUDA-graph replay) the result must equal the launch without the flag bit for bit -- including lengths whose
    new row is the first / last row of a tile, where the tile holding it must NOT be fetched early."""
    L, H, cap, Hq, hidden = 2, 4, 6144, 28, 3584
    g = torch.Generator().manual_seed(kv_after)
    w = (torch.randn((Hq + 2 * H) * 128, hidden, generator=g) / math.sqrt(hidden)).to(torch.bfloat16).to(DEV)
    b = (torch.randn((Hq + 2 * H) * 128, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    nw = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16).to(DEV)
    cos, sin = (t.to(DEV) for t in tables)
    row = torch.tensor([kv_after - 1], dtype=torch.int32, device=DEV)
    after = torch.tensor([kv_after], dtype=torch.int32, device=DEV)
    pos = torch.tensor([1234], dtype=torch.int32, device=DEV)
    outs = []
    for early in (False, True):
        _, arena, view = make_arena(abi, L, H, cap, seed=7)
        q_rot = torch.zeros(1, Hq, 128, dtype=torch.bfloat16, device=DEV)
        o = torch.zeros(1, Hq, 128, dtype=torch.bfloat16, device=DEV)
        ws = torch.zeros(abi.attn_workspace_bytes(1, Hq, 128, splits) // 4 + 16, dtype=torch.float32, device=DEV) if splits > 1 else None
        xs = [torch.randn(hidden, generator=torch.Generator().manual_seed(100 + i)).to(torch.bfloat16).to(DEV) for i in range(3)]

        def chain(x):
            abi.gemv_qkv_rope_append(view, 0, 1, w, x, nw, 1e-6, b, Hq, q_rot, row, pos, cos, sin)
            abi.attn_causal(view, 0, 1, q_rot, o, 1, Hq, cap, 1.0 / math.sqrt(128), splits, ws, kv_len_dev=after, early_kv=early)
        chain(xs[0])
        torch.cuda.synchronize()
        graphs = []
        for x in xs:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                chain(x)
            graphs.append(gr)
        res = []
        for _ in range(3):                      # every replay rewrites row kv_after - 1 with another token's K/V just before attending
            for gr in graphs:
                gr.replay()
                res.append(o.clone())
        torch.cuda.synchronize()
        outs.append(torch.stack(res).cpu())
    assert torch.equal(bits(outs[0]), bits(outs[1]))
    assert not torch.equal(bits(outs[0][0]), bits(outs[0][1]))          # the appended row matters (different tokens, different outputs)


@pytest.mark.parametrize("t,n,k", [(8, 152064, 3584), (16, 1000, 512), (3, 130, 64)])
def test_gemm_tokens_lm_head_argmax_and_streams_advance(abi, t, n, k):
    """Final norm + lm_head + suppress mask + argmax in the projection's epilogue == torch.argmax over the masked fp32 view of the bf16
    logits (first index of the maximum, ties included: a few logits are made exactly equal); svlm_streams_advance hands the ids to the
    ACTIVE streams only and advances their counters."""
    g = torch.Generator().manual_seed(n + t)
    x = torch.randn(t, k, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(n, k, generator=g) * 0.05).to(torch.bfloat16).to(DEV)
    w[n // 2] = w[n // 3]                                            # two identical rows: an exact tie somewhere in the middle
    nw = (1 + 0.1 * torch.randn(k, generator=g)).to(torch.bfloat16).to(DEV)
    sup = torch.zeros(n, dtype=torch.bool)
    sup[torch.randperm(n, generator=g)[: n // 10]] = True
    h = torch.empty_like(x)
    abi.add_rmsnorm(x, None, nw, None, h, 1e-6)
    logits = torch.empty(t, n, dtype=torch.bfloat16, device=DEV)
    abi.gemm_tokens(h, w, logits)
    masked = logits.float().masked_fill(sup.to(DEV), float("-inf"))
    want = torch.argmax(masked, dim=-1)
    # force a tie for the winner of token 0 at a LOWER index that is not suppressed: the first index must win
    keys = torch.zeros(t, dtype=torch.int64, device=DEV)
    abi.gemm_tokens_lm_head_argmax(x, w, keys, norm_weight=nw, eps=1e-6, suppress_u8=sup.to(torch.uint8).to(DEV))
    got = (0xFFFFFFFF - (keys & 0xFFFFFFFF)).cpu()
    assert torch.equal(got, want.cpu()), (got.tolist(), want.cpu().tolist())
    tok = torch.full((t,), 7, dtype=torch.int64, device=DEV)
    active = torch.tensor([b % 3 != 1 for b in range(t)], dtype=torch.int32, device=DEV)
    kv_len = torch.arange(100, 100 + t, dtype=torch.int32, device=DEV)
    kv_after = torch.zeros(t, dtype=torch.int32, device=DEV)
    nxt = torch.arange(500, 500 + t, dtype=torch.int32, device=DEV)
    abi.streams_advance(keys, tok, active, kv_len, kv_after, nxt)
    a = active.bool().cpu()
    assert torch.equal(tok.cpu(), torch.where(a, want.cpu(), torch.full((t,), 7)))
    assert torch.equal(kv_len.cpu(), (torch.arange(100, 100 + t) + a.int()).int())
    assert torch.equal(kv_after.cpu(), kv_len.cpu() + 1) and torch.equal(nxt.cpu(), (torch.arange(500, 500 + t) + a.int()).int())
    assert int(keys.abs().sum()) == 0
