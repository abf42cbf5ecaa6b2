"""Per-session side channel between the session loop and the patched forwards.

Same fields and meaning as the reference's `inference/streaming_args.py:1-10`
(`pos_mode`, `all_text`, `input_ids`, `video_grid_thw`, `second_per_grid_ts`; append mode
adds `last_cache_position`, `qwen2_5/model_forward.py:153`).  Both position modes run on the
B200 path (patched forwards and the session-runner fast path behind `generate`).
"""


class StreamingArgs:
    def __init__(self, pos_mode: str, all_text: bool = False):
        assert pos_mode in ["append", "shrink"], "pos_mode must be in ['append', 'shrink']"
        self.pos_mode = pos_mode
        self.all_text = all_text
        self.input_ids = None            # full ids of the current pruned sequence (shrink mode)
        self.video_grid_thw = None
        self.second_per_grid_ts = None
        self.last_cache_position = None
