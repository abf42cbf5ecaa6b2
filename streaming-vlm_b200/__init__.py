"""streaming-vlm_b200: B200-native (sm_100a) streaming-inference hot path of StreamingVLM.

Layout (only what the path needs):
    csrc/                 CUDA kernels + the C-ABI (`libsvlm_b200.so`, header `include/svlm_b200.h`)
    _abi.py               ctypes binding of the C-ABI + `torch.library` custom ops (`svlm::*`)
    config.py             Qwen2.5-VL geometry presets (tiny / 7B / 32B)
    utils/get_qwen_range  token-pattern range finder      (reference utils/get_qwen_range.py)
    inference/            the reference's host-side interface for this path:
        streaming_args.py        StreamingArgs             (reference inference/streaming_args.py)
        generate/streaming_cache StreamingCache (arena)    (reference generate/streaming_cache.py)
        inference.py             prune / process_past_kv   (reference inference/inference.py:50-172)
        qwen2_5/                 patched forwards + `convert_qwen2_5_to_streaming`
    engine.py             per-stream session runner (evict -> ViT -> prefill -> decode)
    synthetic.py          synthetic token-stream grammar and frames (no tokenizer / video offline)

There is no CPU fallback: every op raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
