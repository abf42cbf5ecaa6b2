"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Frame ingest: uint8 RGB frames -> ViT patch rows.

Restates what the reference's `processor(text=..., videos=clip)` call (`inference/inference.py:390-395`) does to a clip
that is already at its target size (the resize of `livecc_utils/.../video_process_patch.py:140-153` is a no-op for frames
whose sides are multiples of 28 inside the pixel budget, which is what the synthetic 448 x 448 frames are):

* rescale + normalise, fused the way HF's `BaseImageProcessorFast.rescale_and_normalize` does it
  (`transformers/image_processing_utils_fast.py`, 5.5.0): `mean' = fp32(mean) * (1 / rescale)`, `std' = fp32(std) *
  (1 / rescale)` with rescale = 1/255, then `(x - mean') / std'` in fp32 (torchvision `normalize`: sub, then div);
* patch layout of `Qwen2VLVideoProcessor._preprocess` (`models/qwen2_vl/video_processing_qwen2_vl.py`):
  `(gt, Tp, C, gh/m, m, P, gw/m, m, P) -> (gt, gh/m, gw/m, m, m, C, Tp, P, P)`, flattened to `[gt*gh*gw, C*Tp*P*P]`.

Pinned: `tests/golden/frame_ingest.npz` holds the output of HF 5.5's `Qwen2VLVideoProcessor` itself on seeded frames
(`tests/golden/make_golden.py`); `tests/test_oracle_golden.py` checks this restatement against it bit for bit in fp32.
The reference pins transformers 4.52.4 with `use_fast=False` (numpy); the two agree to the last fp32 bit or one ulp, and
the ViT consumes the values as bf16.
"""
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def normalise_constants():
    """(mean', std') as fp32 tensors [3], exactly as HF fuses them."""
    inv_rescale = 1.0 / (1.0 / 255.0)
    return torch.tensor(CLIP_MEAN) * inv_rescale, torch.tensor(CLIP_STD) * inv_rescale


def frames_to_patches(frames_thwc: torch.Tensor, patch: int = 14, temporal_patch: int = 2, merge: int = 2):
    """uint8 [T, H, W, 3] (decord's layout) -> (fp32 [gt*gh*gw, 3*Tp*P*P], (gt, gh, gw))."""
    assert frames_thwc.dtype == torch.uint8 and frames_thwc.dim() == 4 and frames_thwc.shape[-1] == 3
    t, h, w, c = frames_thwc.shape
    assert t % temporal_patch == 0 and h % (patch * merge) == 0 and w % (patch * merge) == 0
    mean, std = normalise_constants()
    x = frames_thwc.permute(0, 3, 1, 2).to(torch.float32)
    x = (x - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    gt, gh, gw = t // temporal_patch, h // patch, w // patch
    x = x.reshape(gt, temporal_patch, c, gh // merge, merge, patch, gw // merge, merge, patch).permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return x.reshape(gt * gh * gw, c * temporal_patch * patch * patch).contiguous(), (gt, gh, gw)


# ---------------------------------------------------------------------------------------------------------------------
# Spatial resize of a clip: `livecc_utils/.../video_process_patch.py:133-153` (`_spatial_resize_video`): target size from
# qwen_vl_utils' `smart_resize` (third-party, pinned 0.0.11 by the reference; the same function ships in transformers as
# `models.qwen2_vl.image_processing_qwen2_vl.smart_resize`, against which the restatement below is tested), then
# `torchvision.transforms.functional.resize(video, [h, w], BICUBIC, antialias=True)` on the uint8 TCHW clip.
# torchvision runs ATen's separable anti-aliased bicubic (`_upsample_bicubic2d_aa`: Keys kernel a = -0.5, support 2 * max(scale, 1),
# weights normalised per output pixel, horizontal pass first, fp32 intermediate) and rounds half-to-even back to uint8.
# ---------------------------------------------------------------------------------------------------------------------
import math

import numpy as np


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280):
    """Both sides multiples of `factor`, pixel count inside [min_pixels, max_pixels], aspect ratio kept as far as possible."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
    h_bar = max(factor, round(height / factor) * factor)
    w_bar = max(factor, round(width / factor) * factor)
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def aa_bicubic_taps(n_in: int, n_out: int):
    """Per output index: first input index, tap count, and the normalised fp32 weights (ATen `_compute_indices_min_size_weights_aa`
    with the bicubic filter, align_corners = False).  Returns (start int32 [n_out], count int32 [n_out], weights fp32 [n_out, kmax])."""
    scale = np.float32(n_in) / np.float32(n_out)
    support = np.float32(2.0) * scale if scale >= 1.0 else np.float32(2.0)
    invscale = np.float32(1.0) / scale if scale >= 1.0 else np.float32(1.0)
    kmax = int(math.ceil(float(support))) * 2 + 1
    start = np.zeros(n_out, np.int32)
    count = np.zeros(n_out, np.int32)
    weights = np.zeros((n_out, kmax), np.float32)
    a = np.float32(-0.5)
    for i in range(n_out):
        center = scale * np.float32(i + 0.5)
        xmin = max(int(center - support + np.float32(0.5)), 0)
        xsize = min(int(center + support + np.float32(0.5)), n_in) - xmin
        xsize = max(min(xsize, kmax), 0)
        w = np.zeros(kmax, np.float32)
        total = np.float32(0.0)
        for j in range(xsize):
            x = abs((np.float32(j + xmin) - center + np.float32(0.5)) * invscale)
            if x < 1.0:
                v = ((a + np.float32(2.0)) * x - (a + np.float32(3.0))) * x * x + np.float32(1.0)
            elif x < 2.0:
                v = (((x - np.float32(5.0)) * x + np.float32(8.0)) * x - np.float32(4.0)) * a
            else:
                v = np.float32(0.0)
            w[j] = np.float32(v)
            total = np.float32(total + w[j])
        if total != 0:
            w[:xsize] = w[:xsize] / total
        start[i], count[i], weights[i] = xmin, xsize, w
    return start, count, weights


def resize_bicubic_aa(frames_thwc: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """uint8 [T, H, W, 3] -> uint8 [T, out_h, out_w, 3]: horizontal pass into fp32, vertical pass, round half to even, clamp.
    Accumulation is sequential fp32 `acc = acc + w * v` (separate multiply and add roundings), which is what the CUDA kernel does."""
    x = frames_thwc.numpy().astype(np.float32)
    t, h, w, c = x.shape
    sx, cx, wx = aa_bicubic_taps(w, out_w)
    sy, cy, wy = aa_bicubic_taps(h, out_h)
    tmp = np.zeros((t, h, out_w, c), np.float32)
    for j in range(wx.shape[1]):
        idx = np.minimum(sx + j, w - 1)
        live = (j < cx).astype(np.float32)[None, None, :, None]
        tmp = tmp + (wx[:, j][None, None, :, None] * x[:, :, idx, :]).astype(np.float32) * live
    out = np.zeros((t, out_h, out_w, c), np.float32)
    for j in range(wy.shape[1]):
        idx = np.minimum(sy + j, h - 1)
        live = (j < cy).astype(np.float32)[None, :, None, None]
        out = out + (wy[:, j][None, :, None, None] * tmp[:, idx, :, :]).astype(np.float32) * live
    return torch.from_numpy(np.clip(np.rint(out), 0, 255).astype(np.uint8))
