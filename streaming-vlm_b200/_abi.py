"""ctypes binding of the C ABI (`include/svlm_b200.h`) and the `svlm::*` torch custom ops.

The shared library is built in-tree by `build()` (plain `nvcc -shared`, sm_100a only, no torch
headers) and loaded with ctypes; tensors cross the boundary as raw device pointers plus the
current CUDA stream.  There is no fallback: if the library is missing or an entry point returns
an error code, a RuntimeError with the library's message is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libsvlm_b200.so")
SOURCES = ["abi.cu", "kv_ops.cu", "attn_fwd.cu", "attn_decode.cu", "gemm.cu", "gemm_tokens.cu", "gemv.cu"]
HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "svlm_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-rdc=false",
              "-Xcompiler", "-fPIC", "-shared"]

EXPORTS = ["svlm_version", "svlm_last_error", "svlm_kv_compact_rope", "svlm_mrope_positions", "svlm_qkv_rope_append",
           "svlm_attn_workspace_bytes", "svlm_attn_causal", "svlm_vit_rope", "svlm_vit_attn", "svlm_gemm_bf16", "svlm_gemm_bf16_ex", "svlm_gemm_tokens", "svlm_gemm_tokens_qkv_append", "svlm_gemm_tokens_lm_head_argmax", "svlm_streams_advance",
           "svlm_rmsnorm", "svlm_add_rmsnorm", "svlm_swiglu", "svlm_gemv_bf16", "svlm_gemv_qkv_rope_append", "svlm_lm_head_argmax", "svlm_lm_head_sample", "svlm_mark_seen", "svlm_embed_tokens", "svlm_prefetch_l2", "svlm_find_segment", "svlm_qkv_rope_append_streams", "svlm_attn_decode_streams",
           "svlm_frames_to_patches", "svlm_frames_resize_bicubic", "svlm_tp_gemv_allreduce", "svlm_tp_allreduce_rows",
           "svlm_kv_compact_rope_host"]

MAX_RUNS = 32


class Arena(Structure):
    _fields_ = [("base", c_void_p), ("layers", c_int32), ("kv_heads", c_int32), ("head_dim", c_int32),
                ("capacity", c_int32)]


class StreamSlot(Structure):
    """`svlm_stream_slot_t` (include/svlm_b200.h section 8)."""
    _fields_ = [("arena_base", c_void_p), ("buffer", c_int32), ("row_dev", c_void_p), ("pos_dev", c_void_p), ("kv_len_dev", c_void_p)]


MAX_STREAMS = 16
MAX_TP = 8
TP_DECODE_OFFSET = 1024


class TpCtx(Structure):
    """`svlm_tp_ctx_t` (include/svlm_b200.h section 10)."""
    _fields_ = [("world", c_int32), ("rank", c_int32), ("peer_base", c_void_p * MAX_TP), ("region_bytes", ctypes.c_uint64),
                ("seq", c_void_p), ("done", c_void_p)]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, s) for s in SOURCES] + [os.path.join(_CSRC, "common.cuh"), os.path.join(_CSRC, "attn_common.cuh"), HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libsvlm_b200.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    r = subprocess.run(cmd, cwd=_CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library (no compute is attempted here, so this also works on a CPU-only box)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first. "
                           "There is no CPU / PyTorch fallback for this path.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32p = c_void_p, POINTER(c_int32)
    L.svlm_version.restype = c_int
    L.svlm_last_error.restype = c_char_p
    L.svlm_kv_compact_rope.argtypes = [POINTER(Arena), c_int, i32p, c_int, vp, c_int, vp, vp, c_int, i32p, vp]
    L.svlm_mrope_positions.argtypes = [vp, c_int, vp, c_int, c_int, c_int, c_int64, vp, c_int, vp, vp, vp]
    L.svlm_qkv_rope_append.argtypes = [POINTER(Arena), c_int, c_int, vp, vp, vp, c_int64, c_int64, c_int64, c_int, c_int, vp,
                                       c_int, vp, vp, c_int, vp, vp, vp, c_int, i32p, vp]
    L.svlm_attn_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.svlm_attn_workspace_bytes.restype = c_size_t
    L.svlm_attn_causal.argtypes = [POINTER(Arena), c_int, c_int, vp, vp, c_int, c_int, c_int, vp, c_int, vp, c_size_t, c_float, c_int, vp]
    L.svlm_vit_rope.argtypes = [vp, vp, vp, c_int, c_int, c_int, vp]
    L.svlm_vit_attn.argtypes = [vp, vp, c_int, vp, c_int, c_int, c_int, c_float, vp]
    L.svlm_gemm_bf16.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp, vp]
    L.svlm_gemm_bf16_ex.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp, c_int, vp]
    L.svlm_gemm_tokens.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, c_int, vp, c_float, vp, c_int, vp, vp]
    L.svlm_gemm_tokens_qkv_append.argtypes = [POINTER(Arena), POINTER(StreamSlot), c_int, c_int, vp, vp, vp, vp, c_float, vp, c_int, c_int, c_int,
                                              vp, vp, vp, c_int, c_int, vp]
    L.svlm_gemm_tokens_lm_head_argmax.argtypes = [vp, vp, vp, c_float, vp, c_int, vp, c_int, c_int, c_int, vp, vp]
    L.svlm_streams_advance.argtypes = [vp, vp, vp, vp, vp, vp, c_int, vp]
    L.svlm_rmsnorm.argtypes = [vp, vp, vp, c_int, c_int, c_float, vp]
    L.svlm_add_rmsnorm.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_float, vp]
    L.svlm_swiglu.argtypes = [vp, vp, c_int, c_int, c_int64, c_int64, vp]
    L.svlm_gemv_bf16.argtypes = [vp, vp, vp, c_float, vp, vp, vp, c_int, c_int, c_int, vp]
    L.svlm_gemv_qkv_rope_append.argtypes = [POINTER(Arena), c_int, c_int, vp, vp, vp, c_float, vp, c_int, c_int, vp, vp, vp, vp, vp, c_int, vp]
    L.svlm_lm_head_argmax.argtypes = [vp, vp, vp, c_float, vp, c_int, c_int, vp, vp, vp, vp, vp, vp]
    L.svlm_lm_head_sample.argtypes = [vp, vp, vp, c_float, vp, c_int, c_int, vp, vp, vp, vp, vp, vp, ctypes.c_double, ctypes.c_double, c_int, vp, vp, vp]
    L.svlm_mark_seen.argtypes = [vp, c_int, vp, c_int64, c_int, vp]
    L.svlm_embed_tokens.argtypes = [vp, c_int, c_int64, vp, c_int64, vp, c_int, c_int, vp, vp, vp]
    L.svlm_prefetch_l2.argtypes = [vp, c_size_t, vp]
    L.svlm_find_segment.argtypes = [vp, c_int64, POINTER(c_int64), c_int, c_int64, c_int, c_int, c_int64, POINTER(c_int64), POINTER(c_int64),
                                    POINTER(c_int64)]
    L.svlm_tp_gemv_allreduce.argtypes = [POINTER(TpCtx), c_int, vp, vp, vp, vp, c_int, c_int, vp]
    L.svlm_tp_allreduce_rows.argtypes = [POINTER(TpCtx), c_int, vp, vp, c_int, c_int, vp]
    L.svlm_frames_resize_bicubic.argtypes = [vp, c_int, c_int, c_int, vp, c_int, c_int, vp, vp, vp, vp, c_int, vp, vp, vp, c_int, vp]
    L.svlm_frames_to_patches.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), vp, c_int, vp]
    L.svlm_qkv_rope_append_streams.argtypes = [POINTER(Arena), POINTER(StreamSlot), c_int, c_int, vp, vp, vp, c_int64, c_int64, c_int64, c_int,
                                               vp, vp, vp, c_int, i32p, vp]
    L.svlm_attn_decode_streams.argtypes = [POINTER(Arena), POINTER(StreamSlot), c_int, c_int, vp, vp, c_int, c_int, c_int, vp, c_size_t,
                                           c_float, c_int, vp]
    L.svlm_kv_compact_rope_host.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, i32p, c_int, i32p, c_int, vp, vp, c_int, i32p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("svlm_last_error", "svlm_attn_workspace_bytes", "svlm_version"):
            fn.restype = c_int
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().svlm_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("svlm ops run on CUDA tensors only (no CPU fallback)")


def i32_array(vals):
    return (c_int32 * len(vals))(*[int(v) for v in vals])


class ArenaView:
    """A torch bf16 tensor [2, 3, L, H, cap, D] described for the C ABI."""

    def __init__(self, tensor: torch.Tensor):
        assert tensor.dtype == torch.bfloat16 and tensor.dim() == 6 and tensor.shape[0] == 2 and tensor.shape[1] == 3
        assert tensor.is_contiguous()
        _req_cuda(tensor)
        self.tensor = tensor
        _, _, self.layers, self.kv_heads, self.capacity, self.head_dim = tensor.shape
        self.c = Arena(tensor.data_ptr(), self.layers, self.kv_heads, self.head_dim, self.capacity)
        self.ref = ctypes.byref(self.c)


# ------------------------------------------------------------------------------------------------
# Thin functional wrappers (raw fast path used by the engine)
# ------------------------------------------------------------------------------------------------
def kv_compact_rope(arena: ArenaView, src_buffer: int, runs, pos: torch.Tensor, cos_table, sin_table, section):
    """runs: list of (src_start, dst_start, length)."""
    _req_cuda(pos, cos_table, sin_table)
    flat = [x for r in runs for x in r]
    check(lib().svlm_kv_compact_rope(arena.ref, src_buffer, i32_array(flat), len(runs), pos.data_ptr(), pos.stride(0),
                                     cos_table.data_ptr(), sin_table.data_ptr(), cos_table.shape[0], i32_array(section),
                                     _stream()), "svlm_kv_compact_rope")


def mrope_positions(ids: torch.Tensor, n: int, grids: torch.Tensor, n_grids: int, pos: torch.Tensor, next_pos: torch.Tensor,
                    status: torch.Tensor, spatial_merge: int = 2, t_step: int = 2, video_token_id: int = 151656):
    _req_cuda(ids, grids, pos, next_pos, status)
    assert ids.dtype == torch.int64 and pos.dtype == torch.int32 and grids.dtype == torch.int32
    check(lib().svlm_mrope_positions(ids.data_ptr(), n, grids.data_ptr(), n_grids, spatial_merge, t_step, video_token_id,
                                     pos.data_ptr(), pos.stride(0), next_pos.data_ptr(), status.data_ptr(), _stream()),
          "svlm_mrope_positions")


def qkv_rope_append(arena: ArenaView, buffer: int, layer: int, q, k, v, n_tokens: int, q_heads: int, q_out, start: int, pos,
                    cos_table, sin_table, section, start_dev=None, text_pos_dev=None):
    """q_heads == 0 (q = q_out = None): append-only, K_raw / K_rot / V rows of the new tokens."""
    _req_cuda(q, k, v, q_out)
    check(lib().svlm_qkv_rope_append(arena.ref, buffer, layer, _ptr(q), k.data_ptr(), v.data_ptr(), 0 if q is None else q.stride(0), k.stride(0),
                                     v.stride(0), n_tokens, q_heads, _ptr(q_out), start, _ptr(start_dev), _ptr(pos),
                                     0 if pos is None else pos.stride(0), _ptr(text_pos_dev), cos_table.data_ptr(),
                                     sin_table.data_ptr(), cos_table.shape[0], i32_array(section), _stream()),
          "svlm_qkv_rope_append")


def attn_workspace_bytes(n_q_tokens: int, q_heads: int, head_dim: int, num_splits: int) -> int:
    return int(lib().svlm_attn_workspace_bytes(n_q_tokens, q_heads, head_dim, num_splits))


ATTN_EARLY_KV = 1


def attn_causal(arena: ArenaView, buffer: int, layer: int, q, out, n_q_tokens: int, q_heads: int, kv_len: int, scale: float,
                num_splits: int = 1, workspace=None, kv_len_dev=None, early_kv: bool = False):
    """`early_kv` (decode, with `kv_len_dev`): the device-side length and the cache rows below the one being appended were not written
    by the launch right before this one -- the kernel may fetch them before its grid dependency resolves (SVLM_ATTN_EARLY_KV)."""
    _req_cuda(q, out)
    check(lib().svlm_attn_causal(arena.ref, buffer, layer, q.data_ptr(), out.data_ptr(), n_q_tokens, q_heads, kv_len,
                                 _ptr(kv_len_dev), num_splits, _ptr(workspace),
                                 0 if workspace is None else workspace.numel() * workspace.element_size(), scale,
                                 ATTN_EARLY_KV if early_kv else 0, _stream()),
          "svlm_attn_causal")


def vit_rope(qkv, cos, sin, n_tokens: int, heads: int, head_dim: int):
    _req_cuda(qkv, cos, sin)
    assert cos.dtype == torch.float32 and sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
    check(lib().svlm_vit_rope(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), n_tokens, heads, head_dim, _stream()), "svlm_vit_rope")


def vit_attn(qkv, cu_seqlens, out, n_tokens: int, heads: int, head_dim: int, scale: float):
    _req_cuda(qkv, cu_seqlens, out)
    assert cu_seqlens.dtype == torch.int32
    check(lib().svlm_vit_attn(qkv.data_ptr(), cu_seqlens.data_ptr(), cu_seqlens.numel() - 1, out.data_ptr(), n_tokens, heads,
                              head_dim, scale, _stream()), "svlm_vit_attn")


def gemm_bf16(a, w, bias, out, act: int = 0, out_rows=None):
    _req_cuda(a, w, out)
    m, k = a.shape
    n = w.shape[0]
    assert w.shape[1] == k and a.is_contiguous() and w.is_contiguous() and out.is_contiguous()
    check(lib().svlm_gemm_bf16(a.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(), m, n, k, act, _ptr(out_rows), _stream()),
          "svlm_gemm_bf16")


EPI_NONE, EPI_GELU, EPI_RESIDUAL, EPI_SWIGLU = 0, 1, 2, 3


def gemm_ex(a, w, out, bias=None, residual=None, epilogue: int = EPI_NONE, out_rows=None, tile_n: int = 0):
    """out = epilogue(a @ w.T + bias) on the persistent tcgen05 GEMM (see include/svlm_b200.h, svlm_gemm_bf16_ex)."""
    _req_cuda(a, w, out, bias, residual)
    m, k = a.shape
    n = w.shape[0]
    assert w.shape[1] == k and a.is_contiguous() and w.is_contiguous() and out.is_contiguous()
    assert out.shape[1] == (n // 2 if epilogue == EPI_SWIGLU else n)
    check(lib().svlm_gemm_bf16_ex(a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(), m, n, k, epilogue, _ptr(out_rows),
                                  tile_n, _stream()), "svlm_gemm_bf16_ex")


def gemm_tokens(x, w, out, bias=None, residual=None, epilogue: int = EPI_NONE, splits: int = 0, norm_weight=None, eps: float = 1e-6,
                stats_in=None, stats_out=None):
    """out[T, N] = epilogue(norm(x)[T, K] @ w.T + bias) for T <= 64 tokens: the weight-streaming swap-AB GEMM of the batched decode step
    (include/svlm_b200.h, svlm_gemm_tokens).  epilogue EPI_SWIGLU takes w = [gate rows ; up rows] and writes [T, N / 2];
    `norm_weight`: RMSNorm of x fused into the kernel; `stats_out` (fp32 [ceil(N / 128), T]): per-tile sums of squares of the rows
    written, `stats_in`: such a table for x, left by the launch that produced it."""
    _req_cuda(x, w, out, bias, residual, norm_weight, stats_in, stats_out)
    assert stats_in is None or (stats_in.dtype == torch.float32 and stats_in.is_contiguous() and stats_in.shape[1] == x.shape[0])
    assert stats_out is None or (stats_out.dtype == torch.float32 and stats_out.is_contiguous()
                                 and tuple(stats_out.shape) == (-(-w.shape[0] // 128), x.shape[0]))
    t, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k and x.is_contiguous() and w.is_contiguous() and out.is_contiguous()
    assert tuple(out.shape) == (t, n // 2 if epilogue == EPI_SWIGLU else n)
    assert residual is None or (residual.is_contiguous() and residual.shape == out.shape)
    check(lib().svlm_gemm_tokens(x.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(), t, n, k, epilogue, splits,
                                 _ptr(norm_weight), eps, _ptr(stats_in), 0 if stats_in is None else stats_in.shape[0], _ptr(stats_out),
                                 _stream()), "svlm_gemm_tokens")


def gemm_tokens_qkv_append(geometry: "ArenaView", slots, layer, x, w, bias, norm_weight, eps, q_heads, q_rot, cos_table, sin_table,
                           splits: int = 0, stats_in=None):
    """Batched decode step: RMSNorm + q/k/v projection + bias + mRoPE + arena append of token b into stream b, one launch
    (include/svlm_b200.h, svlm_gemm_tokens_qkv_append)."""
    _req_cuda(x, w, bias, norm_weight, q_rot, stats_in)
    assert x.is_contiguous() and w.is_contiguous() and q_rot.is_contiguous() and x.shape[0] == len(slots)
    assert stats_in is None or (stats_in.dtype == torch.float32 and stats_in.is_contiguous() and stats_in.shape[1] == x.shape[0])
    check(lib().svlm_gemm_tokens_qkv_append(geometry.ref, slots, len(slots), layer, x.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(norm_weight),
                                            eps, _ptr(stats_in), 0 if stats_in is None else stats_in.shape[0], q_heads, x.shape[1],
                                            q_rot.data_ptr(), cos_table.data_ptr(), sin_table.data_ptr(),
                                            cos_table.shape[0], splits, _stream()), "svlm_gemm_tokens_qkv_append")


def gemm_tokens_lm_head_argmax(x, w, keys, norm_weight=None, eps: float = 1e-6, stats_in=None, suppress_u8=None):
    """keys[t] = packed (logit, ~row) maximum over the non-suppressed rows of norm(x)[t] @ w.T: the batched lm_head + argmax without a
    logits tensor (include/svlm_b200.h, svlm_gemm_tokens_lm_head_argmax).  `keys`: int64 / uint64 [T], zero on entry."""
    _req_cuda(x, w, keys, norm_weight, stats_in, suppress_u8)
    t, k = x.shape
    assert w.shape[1] == k and x.is_contiguous() and w.is_contiguous() and keys.numel() >= t and keys.element_size() == 8
    assert suppress_u8 is None or (suppress_u8.dtype == torch.uint8 and suppress_u8.numel() == w.shape[0])
    check(lib().svlm_gemm_tokens_lm_head_argmax(x.data_ptr(), w.data_ptr(), _ptr(norm_weight), eps, _ptr(stats_in),
                                                0 if stats_in is None else stats_in.shape[0], _ptr(suppress_u8), t, w.shape[0], k,
                                                keys.data_ptr(), _stream()), "svlm_gemm_tokens_lm_head_argmax")


def streams_advance(keys, tok, active, kv_len, kv_after, next_pos):
    _req_cuda(keys, tok, active, kv_len, kv_after, next_pos)
    assert tok.dtype == torch.int64 and all(t_.dtype == torch.int32 for t_ in (active, kv_len, kv_after, next_pos))
    check(lib().svlm_streams_advance(keys.data_ptr(), tok.data_ptr(), active.data_ptr(), kv_len.data_ptr(), kv_after.data_ptr(),
                                     next_pos.data_ptr(), tok.numel(), _stream()), "svlm_streams_advance")


def interleave_gate_up(gate, up, block: int = 64):
    """[gate rows ; up rows] -> rows interleaved in blocks of `block` (the layout the SwiGLU epilogue pairs inside one tile)."""
    n, k = gate.shape
    assert up.shape == gate.shape and n % block == 0
    return torch.stack((gate.reshape(n // block, block, k), up.reshape(n // block, block, k)), dim=1).reshape(2 * n, k).contiguous()


def gemv_bf16(w, x, y, mode: int = 0, norm_weight=None, eps: float = 1e-6, bias=None, residual=None):
    """y[N] = epilogue(W[rows, K] @ prologue(x[K])); mode 0 plain, 1 residual add, 2 SwiGLU (W = [gate; up])."""
    _req_cuda(w, x, y)
    n = y.numel()
    check(lib().svlm_gemv_bf16(w.data_ptr(), x.data_ptr(), _ptr(norm_weight), eps, _ptr(bias), _ptr(residual), y.data_ptr(), n,
                               w.shape[1], mode, _stream()), "svlm_gemv_bf16")


def gemv_qkv_rope_append(arena: ArenaView, buffer: int, layer: int, w, x, norm_weight, eps, bias, q_heads: int, q_rot, row_dev, pos_dev,
                         cos_table, sin_table):
    _req_cuda(w, x, q_rot, row_dev, pos_dev)
    check(lib().svlm_gemv_qkv_rope_append(arena.ref, buffer, layer, w.data_ptr(), x.data_ptr(), norm_weight.data_ptr(), eps, _ptr(bias),
                                          q_heads, w.shape[1], q_rot.data_ptr(), row_dev.data_ptr(), pos_dev.data_ptr(),
                                          cos_table.data_ptr(), sin_table.data_ptr(), cos_table.shape[0], _stream()),
          "svlm_gemv_qkv_rope_append")


def lm_head_argmax(w, x, norm_weight, eps, suppress_u8, key_scratch, token_out, kv_len, kv_after, next_pos):
    _req_cuda(w, x, key_scratch, token_out)
    check(lib().svlm_lm_head_argmax(w.data_ptr(), x.data_ptr(), _ptr(norm_weight), eps, _ptr(suppress_u8), w.shape[0], w.shape[1],
                                    key_scratch.data_ptr(), token_out.data_ptr(), kv_len.data_ptr(), kv_after.data_ptr(),
                                    next_pos.data_ptr(), _stream()), "svlm_lm_head_argmax")


def lm_head_sample(w, x, norm_weight, eps, suppress_u8, key_scratch, token_out, kv_len, kv_after, next_pos, seen=None,
                   repetition_penalty: float = 1.0, temperature: float = 1.0, do_sample: bool = False, rng_state=None, scores_out=None):
    """lm_head with the reference's generation config fused in (repetition penalty over `seen`, temperature, Gumbel-max
    sampling; `rng_state`: int32/uint32 [4] on the device = seed lo, seed hi, step, reserved); see include/svlm_b200.h."""
    _req_cuda(w, x, key_scratch, token_out, seen, rng_state, scores_out)
    check(lib().svlm_lm_head_sample(w.data_ptr(), x.data_ptr(), _ptr(norm_weight), eps, _ptr(suppress_u8), w.shape[0], w.shape[1],
                                    key_scratch.data_ptr(), token_out.data_ptr(), kv_len.data_ptr(), kv_after.data_ptr(),
                                    next_pos.data_ptr(), _ptr(seen), float(repetition_penalty), float(temperature), 1 if do_sample else 0,
                                    _ptr(rng_state), _ptr(scores_out), _stream()), "svlm_lm_head_sample")


def rng_words(seed: int):
    """(seed lo, seed hi, step = 0, 0) as signed 32-bit values for an int32 tensor."""
    def s32(v):
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v >= (1 << 31) else v
    return [s32(seed), s32(seed >> 32), 0, 0]


def mark_seen(ids, n: int, seen, clear_first: bool = True):
    _req_cuda(ids, seen)
    assert ids.dtype == torch.int64 and seen.dtype == torch.uint8
    check(lib().svlm_mark_seen(ids.data_ptr(), n, seen.data_ptr(), seen.numel(), 1 if clear_first else 0, _stream()), "svlm_mark_seen")


def embed_tokens(ids, n: int, embed, video, n_video: int, out, status=None, video_token_id: int = 151656):
    """out[t] = video[rank(t)] where ids[t] is the video placeholder, else embed[ids[t]] (model_forward.py:58, 87-93)."""
    _req_cuda(ids, embed, video, out, status)
    assert ids.dtype == torch.int64 and embed.is_contiguous() and out.is_contiguous() and (video is None or video.is_contiguous())
    check(lib().svlm_embed_tokens(ids.data_ptr(), n, video_token_id, embed.data_ptr(), embed.shape[0], _ptr(video), n_video, embed.shape[1],
                                  out.data_ptr(), _ptr(status), _stream()), "svlm_embed_tokens")


def tp_gemv_allreduce(ctx_ref, which: int, w, x, residual, y):
    """y = residual + bf16(sum over ranks of W_r x_r): row-parallel GEMV fused with the one-shot all-reduce over NVLink peer memory."""
    _req_cuda(w, x, residual, y)
    check(lib().svlm_tp_gemv_allreduce(ctx_ref, which, w.data_ptr(), x.data_ptr(), residual.data_ptr(), y.data_ptr(), y.numel(), w.shape[1],
                                       _stream()), "svlm_tp_gemv_allreduce")


def tp_allreduce_rows(ctx_ref, which: int, residual, out, rows: int, cols: int):
    """out = residual + bf16(sum over ranks of the bf16 partial products sitting in region `which` of the symmetric buffer)."""
    _req_cuda(residual, out)
    check(lib().svlm_tp_allreduce_rows(ctx_ref, which, residual.data_ptr(), out.data_ptr(), rows, cols, _stream()), "svlm_tp_allreduce_rows")


def find_segment(ids_np, pattern, end_token: int, index: int, contain_lf: bool, lf_token: int):
    """(start, end) of the `index`-th non-overlapping segment `pattern ... end_token` in a contiguous int64 numpy array: the native
    scan behind `get_qwen_range` (include/svlm_b200.h, svlm_find_segment; host only, no GPU needed).  IndexError when out of range."""
    pat = (c_int64 * len(pattern))(*pattern)
    s, e = c_int64(0), c_int64(0)
    rc = lib().svlm_find_segment(ids_np.ctypes.data, ids_np.shape[0], pat, len(pattern), end_token, index, 1 if contain_lf else 0, lf_token,
                                 ctypes.byref(s), ctypes.byref(e), None)
    if rc != 0:
        raise IndexError(lib().svlm_last_error().decode(errors="replace"))
    return s.value, e.value


def make_slots(caches, row_devs, pos_devs, kv_len_devs):
    """HOST array of `svlm_stream_slot_t`, one per stream: arena + live buffer of `caches[b]`, and the device scalars that name
    the row / position of the token being appended and the attention bound."""
    n = len(caches)
    if not 1 <= n <= MAX_STREAMS:
        raise ValueError(f"1..{MAX_STREAMS} streams per launch, got {n}")
    arr = (StreamSlot * n)()
    for b, c in enumerate(caches):
        arr[b].arena_base = c.arena.data_ptr()
        arr[b].buffer = c.cur
        arr[b].row_dev = row_devs[b].data_ptr()
        arr[b].pos_dev = pos_devs[b].data_ptr()
        arr[b].kv_len_dev = kv_len_devs[b].data_ptr()
    return arr


def qkv_rope_append_streams(geometry: "ArenaView", slots, layer, q, k, v, q_heads, q_out, cos_table, sin_table, section):
    """Decode step of len(slots) streams: row b of q/k/v goes to stream b (see include/svlm_b200.h section 8)."""
    _req_cuda(q, k, v, q_out)
    check(lib().svlm_qkv_rope_append_streams(geometry.ref, slots, len(slots), layer, q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0),
                                             k.stride(0), v.stride(0), q_heads, q_out.data_ptr(), cos_table.data_ptr(),
                                             sin_table.data_ptr(), cos_table.shape[0], i32_array(section), _stream()), "svlm_qkv_rope_append_streams")


def attn_decode_streams(geometry: "ArenaView", slots, layer, q, out, q_heads, max_kv_len, scale, num_splits, workspace, early_kv: bool = False):
    _req_cuda(q, out)
    ws_ptr = workspace.data_ptr() if workspace is not None else None
    ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    check(lib().svlm_attn_decode_streams(geometry.ref, slots, len(slots), layer, q.data_ptr(), out.data_ptr(), q_heads, max_kv_len,
                                         num_splits, ws_ptr, ws_bytes, scale, ATTN_EARLY_KV if early_kv else 0, _stream()),
          "svlm_attn_decode_streams")


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def frames_to_patches(frames_thwc, out=None, patch=14, temporal_patch=2, merge=2, dtype=torch.bfloat16, mean=CLIP_MEAN, std=CLIP_STD):
    """uint8 CUDA frames [T, H, W, 3] -> ViT patch rows [gt*gh*gw, 3*Tp*P*P] (fp32 or bf16) and the (gt, gh, gw) grid."""
    _req_cuda(frames_thwc)
    assert frames_thwc.dtype == torch.uint8 and frames_thwc.dim() == 4 and frames_thwc.shape[-1] == 3 and frames_thwc.is_contiguous()
    t, h, w, _ = frames_thwc.shape
    grid = (t // temporal_patch, h // patch, w // patch)
    if out is None:
        out = torch.empty(grid[0] * grid[1] * grid[2], 3 * temporal_patch * patch * patch, dtype=dtype, device=frames_thwc.device)
    assert out.dtype in (torch.float32, torch.bfloat16) and out.is_contiguous()
    # mean' = mean * 255, std' = std * 255 in fp32, the way HF fuses rescale and normalise
    inv_rescale = 1.0 / (1.0 / 255.0)
    m = (torch.tensor(mean, dtype=torch.float32) * inv_rescale).tolist()
    s_ = (torch.tensor(std, dtype=torch.float32) * inv_rescale).tolist()
    check(lib().svlm_frames_to_patches(frames_thwc.data_ptr(), t, h, w, patch, temporal_patch, merge, (c_float * 3)(*m), (c_float * 3)(*s_),
                                       out.data_ptr(), 1 if out.dtype == torch.bfloat16 else 0, _stream()), "svlm_frames_to_patches")
    return out, grid


_TAP_CACHE = {}


def frames_resize_bicubic(frames_thwc, out_h: int, out_w: int):
    """uint8 CUDA clip [T, H, W, 3] -> uint8 [T, out_h, out_w, 3], anti-aliased bicubic (torchvision `resize(BICUBIC, antialias=True)`)."""
    from .utils.frames import aa_bicubic_taps
    _req_cuda(frames_thwc)
    assert frames_thwc.dtype == torch.uint8 and frames_thwc.dim() == 4 and frames_thwc.shape[-1] == 3 and frames_thwc.is_contiguous()
    t, h, w, _ = frames_thwc.shape
    dev = frames_thwc.device
    key = (h, w, out_h, out_w, str(dev))
    taps = _TAP_CACHE.get(key)
    if taps is None:
        taps = _TAP_CACHE[key] = tuple(torch.from_numpy(a).to(dev) for a in (*aa_bicubic_taps(w, out_w), *aa_bicubic_taps(h, out_h)))
    xs, xc, xw, ys, yc, yw = taps
    tmp = torch.empty(t, h, out_w, 3, dtype=torch.float32, device=dev)
    out = torch.empty(t, out_h, out_w, 3, dtype=torch.uint8, device=dev)
    check(lib().svlm_frames_resize_bicubic(frames_thwc.data_ptr(), t, h, w, out.data_ptr(), out_h, out_w, tmp.data_ptr(), xs.data_ptr(),
                                           xc.data_ptr(), xw.data_ptr(), xw.shape[1], ys.data_ptr(), yc.data_ptr(), yw.data_ptr(),
                                           yw.shape[1], _stream()), "svlm_frames_resize_bicubic")
    return out


def prefetch_l2(t, nbytes=None):
    """Issue L2 prefetches for the first `nbytes` (default: all) of the contiguous tensor `t` on the current stream."""
    _req_cuda(t)
    total = t.numel() * t.element_size()
    n = total if nbytes is None else min(int(nbytes), total)
    check(lib().svlm_prefetch_l2(t.data_ptr(), n, _stream()), "svlm_prefetch_l2")


def add_rmsnorm(a, b, weight, out_sum, out_norm, eps: float):
    """out_sum = a + b (both optional), out_norm = RMSNorm(out_sum) * weight."""
    _req_cuda(a, weight, out_norm)
    rows, cols = a.shape
    check(lib().svlm_add_rmsnorm(a.data_ptr(), _ptr(b), weight.data_ptr(), _ptr(out_sum), out_norm.data_ptr(), rows, cols, eps,
                                 _stream()), "svlm_add_rmsnorm")


def swiglu(gate_up, out):
    _req_cuda(gate_up, out)
    rows, m = out.shape
    assert gate_up.shape[1] == 2 * m and gate_up.stride(1) == 1 and out.stride(1) == 1
    check(lib().svlm_swiglu(gate_up.data_ptr(), out.data_ptr(), rows, m, gate_up.stride(0), out.stride(0), _stream()), "svlm_swiglu")


def rmsnorm(x, weight, out, eps: float):
    _req_cuda(x, weight, out)
    rows, cols = x.shape
    check(lib().svlm_rmsnorm(x.data_ptr(), weight.data_ptr(), out.data_ptr(), rows, cols, eps, _stream()), "svlm_rmsnorm")


# ------------------------------------------------------------------------------------------------
# torch custom ops (`torch.ops.svlm.*`): the thin custom-op layer the patched HF modules call.
# ------------------------------------------------------------------------------------------------
_ops_registered = False


def register_torch_ops():
    global _ops_registered
    if _ops_registered:
        return
    _ops_registered = True
    L = torch.library.Library("svlm", "DEF")
    L.define("kv_compact_rope(Tensor(a!) arena, int src_buffer, int[] runs, Tensor pos, Tensor cos_table, Tensor sin_table, int[] section) -> ()")
    L.define("mrope_positions(Tensor ids, Tensor grids, Tensor(a!) pos, Tensor(b!) next_pos, Tensor(c!) status, int spatial_merge, int t_step, int video_token_id) -> ()")
    L.define("qkv_rope_append(Tensor(a!) arena, int buffer, int layer, Tensor q, Tensor k, Tensor v, Tensor(b!) q_out, int start, Tensor pos, Tensor cos_table, Tensor sin_table, int[] section) -> ()")
    L.define("attn_causal(Tensor arena, int buffer, int layer, Tensor q, int kv_len, float scale, int num_splits) -> Tensor")
    L.define("vit_attn(Tensor(a!) qkv, Tensor cos, Tensor sin, Tensor cu_seqlens, float scale) -> Tensor")
    L.define("gemm_bf16(Tensor a, Tensor w, Tensor? bias, int act, Tensor? out_rows) -> Tensor")
    L.define("rmsnorm(Tensor x, Tensor weight, float eps) -> Tensor")
    L.define("add_rmsnorm(Tensor a, Tensor b, Tensor weight, float eps) -> (Tensor, Tensor)")
    L.define("swiglu(Tensor gate_up) -> Tensor")
    L.define("gemv(Tensor w, Tensor x, int mode, Tensor? norm_weight, float eps, Tensor? bias, Tensor? residual) -> Tensor")
    L.define("frames_to_patches(Tensor frames_thwc, int patch, int temporal_patch, int merge) -> Tensor")
    L.define("frames_resize_bicubic(Tensor frames_thwc, int out_h, int out_w) -> Tensor")

    def _kv_compact_rope(arena, src_buffer, runs, pos, cos_table, sin_table, section):
        kv_compact_rope(ArenaView(arena), src_buffer, [tuple(runs[i:i + 3]) for i in range(0, len(runs), 3)], pos, cos_table,
                        sin_table, section)

    def _mrope_positions(ids, grids, pos, next_pos, status, spatial_merge, t_step, video_token_id):
        mrope_positions(ids.reshape(-1), ids.numel(), grids, grids.shape[0], pos, next_pos, status, spatial_merge, t_step,
                        video_token_id)

    def _qkv_rope_append(arena, buffer, layer, q, k, v, q_out, start, pos, cos_table, sin_table, section):
        qkv_rope_append(ArenaView(arena), buffer, layer, q, k, v, q.shape[0], q_out.shape[1], q_out, start, pos, cos_table,
                        sin_table, section)

    def _attn_causal(arena, buffer, layer, q, kv_len, scale, num_splits):
        out = torch.empty_like(q)
        ws = None
        if num_splits > 1:
            ws = torch.empty(attn_workspace_bytes(q.shape[0], q.shape[1], q.shape[2], num_splits) // 4, dtype=torch.float32,
                             device=q.device)
        attn_causal(ArenaView(arena), buffer, layer, q, out, q.shape[0], q.shape[1], kv_len, scale, num_splits, ws)
        return out

    def _vit_attn(qkv, cos, sin, cu_seqlens, scale):
        n, _, h, d = qkv.shape
        vit_rope(qkv, cos, sin, n, h, d)
        out = torch.empty(n, h, d, dtype=qkv.dtype, device=qkv.device)
        vit_attn(qkv, cu_seqlens, out, n, h, d, scale)
        return out

    def _gemm(a, w, bias, act, out_rows):
        out = torch.empty(a.shape[0], w.shape[0], dtype=a.dtype, device=a.device)
        gemm_bf16(a, w, bias, out, act, out_rows)
        return out

    def _rmsnorm(x, weight, eps):
        out = torch.empty_like(x)
        rmsnorm(x, weight, out, eps)
        return out

    def _add_rmsnorm(a, b, weight, eps):
        s_, h = torch.empty_like(a), torch.empty_like(a)
        add_rmsnorm(a, b, weight, s_, h, eps)
        return s_, h

    def _swiglu(gate_up):
        out = torch.empty(gate_up.shape[0], gate_up.shape[1] // 2, dtype=gate_up.dtype, device=gate_up.device)
        swiglu(gate_up, out)
        return out

    def _gemv(w, x, mode, norm_weight, eps, bias, residual):
        n = w.shape[0] // 2 if mode == 2 else w.shape[0]
        y = torch.empty(n, dtype=x.dtype, device=x.device)
        gemv_bf16(w, x, y, mode, norm_weight, eps, bias, residual)
        return y

    def _frames_to_patches(frames_thwc, patch, temporal_patch, merge):
        return frames_to_patches(frames_thwc, patch=patch, temporal_patch=temporal_patch, merge=merge)[0]

    for name, fn in [("kv_compact_rope", _kv_compact_rope), ("mrope_positions", _mrope_positions),
                     ("qkv_rope_append", _qkv_rope_append), ("attn_causal", _attn_causal), ("vit_attn", _vit_attn),
                     ("gemm_bf16", _gemm), ("rmsnorm", _rmsnorm), ("add_rmsnorm", _add_rmsnorm), ("swiglu", _swiglu), ("gemv", _gemv),
                     ("frames_to_patches", _frames_to_patches), ("frames_resize_bicubic", frames_resize_bicubic)]:
        L.impl(name, fn, "CUDA")
    register_torch_ops._lib = L     # keep the Library object alive
