"""Host side of the frame ingest: target size and tap tables of the spatial resize.

`smart_resize` is qwen_vl_utils' function (third-party dependency of the reference, pinned 0.0.11; the same function ships in
transformers as `models.qwen2_vl.image_processing_qwen2_vl.smart_resize`): the size `livecc_utils/.../video_process_patch.py:
133-146` resizes every clip to.  `aa_bicubic_taps` builds, per output index, the tap window and the normalised weights of ATen's
anti-aliased bicubic filter (`_upsample_bicubic2d_aa`, Keys kernel with a = -0.5), which is what
`torchvision.transforms.functional.resize(..., BICUBIC, antialias=True)` (`:147-152`) runs; `svlm_frames_resize_bicubic` consumes
the tables on the device.
"""
from __future__ import annotations

import math

import numpy as np


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280):
    ratio = max(height, width) / min(height, width)
    if ratio > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {ratio}")
    snap = lambda v, fn: max(factor, fn(v / factor) * factor)
    h2, w2 = snap(height, round), snap(width, round)
    if h2 * w2 > max_pixels:
        shrink = math.sqrt(height * width / max_pixels)
        h2, w2 = snap(height / shrink, math.floor), snap(width / shrink, math.floor)
    elif h2 * w2 < min_pixels:
        grow = math.sqrt(min_pixels / (height * width))
        h2, w2 = math.ceil(height * grow / factor) * factor, math.ceil(width * grow / factor) * factor
    return h2, w2


def aa_bicubic_taps(n_in: int, n_out: int):
    """(start int32 [n_out], count int32 [n_out], weights fp32 [n_out, k]) of the anti-aliased bicubic filter, all in fp32
    arithmetic like ATen's float path (vectorised over the output index)."""
    f32 = np.float32
    scale = f32(n_in) / f32(n_out)
    support = f32(2.0) * scale if scale >= 1.0 else f32(2.0)
    inv = f32(1.0) / scale if scale >= 1.0 else f32(1.0)
    k = int(math.ceil(float(support))) * 2 + 1
    center = scale * (np.arange(n_out, dtype=f32) + f32(0.5))
    start = np.maximum((center - support + f32(0.5)).astype(np.int32), 0)
    count = np.clip(np.minimum((center + support + f32(0.5)).astype(np.int32), n_in) - start, 0, k).astype(np.int32)
    j = np.arange(k, dtype=np.int32)[None, :]
    x = np.abs(((j + start[:, None]).astype(f32) - center[:, None] + f32(0.5)) * inv).astype(f32)
    a = f32(-0.5)
    near = ((a + f32(2.0)) * x - (a + f32(3.0))) * x * x + f32(1.0)
    far = (((x - f32(5.0)) * x + f32(8.0)) * x - f32(4.0)) * a
    w = np.where(x < 1.0, near, np.where(x < 2.0, far, f32(0.0))).astype(f32)
    w = np.where(j < count[:, None], w, f32(0.0)).astype(f32)
    total = np.zeros(n_out, f32)
    for col in range(k):                       # sequential fp32 sum, the order ATen accumulates in
        total = (total + w[:, col]).astype(f32)
    w = np.where(total[:, None] != 0, w / np.where(total == 0, f32(1.0), total)[:, None], w).astype(f32)
    return start.astype(np.int32), count, np.ascontiguousarray(w)
