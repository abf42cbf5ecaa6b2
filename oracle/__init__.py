"""CPU oracle for the StreamingVLM hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (pure torch / numpy, no CUDA) of the reference's
algorithm for the streaming-inference hot path.  Nothing in the product package
(`streaming-vlm_b200/`) imports it; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the checker or as
the timed CPU baseline, never as the thing shipped.

Parity status: PINNED for the integer / bf16-elementwise rows, by execution of the
reference's own functions in the build container (`tests/golden/make_golden.py`
exec's `get_qwen_range`, `prune_id_and_kv_cache`, `process_past_kv`, `get_rope_index`
and `apply_multimodal_rotary_pos_emb` straight from `/root/reference/src` and commits
their outputs as fixtures under `tests/golden/`).  The reference has no tests or golden
vectors of its own.  The frame-ingest rows are pinned against the
libraries the reference calls for them, run in the build container: HF's
`Qwen2VLVideoProcessor` (patch rows, bit-exact), torchvision's anti-aliased bicubic
`resize` (<= 1 LSB), transformers' `smart_resize`; the Qwen2-VL tower against HF's
`Qwen2VisionTransformerPretrainedModel`.  The attention / ViT-attention rows call
flash-attn 2.8.3 in the reference, which has no CPU path: here the oracle is the fp32
softmax(QK^T)V definition, and the CUDA kernels are additionally compared with flash-attn
itself on the GPU box (`tests/test_gpu_kernels.py::test_attn_causal_matches_flash_attn`,
`::test_vit_attn_matches_flash_attn_varlen`).  The `pos_mode="append"` arithmetic
(`model_forward.py:112-153`) is pinned too: the reference's `model_forward` is exec'd with
a stub model and the `position_ids` it hands to the language model are the fixture
(`tests/golden/append_positions.json`).

Module map (reference file:line each one follows is cited in the module docstrings):
    tokens.py      get_qwen_range                      utils/get_qwen_range.py:15-86
    eviction.py    prune / process_past_kv             inference/inference.py:50-172
    cache.py       StreamingCache (list + torch.cat)   generate/streaming_cache.py:6-74
    positions.py   get_rope_index, rotary tables, mRoPE qwen2_5/pos_emb.py:6-188,
                                                       qwen2_5/language_forward.py:11-67
    attention.py   LLM attention (prefill/decode)      qwen2_5/language_forward.py:143-260
    vit.py         ViT encoder / attention / merger    qwen2_5/vision_forward.py:6-101
    model.py       model forward + greedy streaming    qwen2_5/model_forward.py:30-176,
                                                       inference/inference.py:309-517
    frames.py      frame ingest: resize, normalise,    livecc_utils/.../video_process_patch.py:133-153,
                   patch layout                        inference/inference.py:390-395 (HF video processor)
"""
