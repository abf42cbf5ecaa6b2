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
