"""Token-pattern range finder for the Qwen chat grammar (host side, integer-exact).

Mirrors the interface of the reference's `utils/get_qwen_range.py` (`TOKEN_IDS` :2-13,
`SYSTEM_PROMPT_OFFSET` :1, `get_qwen_range` :15-86): closed `(start, end)` interval of the
`index`-th segment of `label` in a flat id sequence, a trailing "\n" absorbed when
`contain_lf`, `user_text` stopping one token before `<|vision_start|>`.

Implementation differs from the reference's O(S * pattern) Python scan: candidate start
and end positions are found with vectorised numpy comparisons and the greedy,
non-overlapping left-to-right pairing is then a walk over those two short sorted arrays
(O(S) vectorised + O(#segments)).  The ids are expected on the HOST (the session keeps
them there), so no device synchronisation happens here.
"""
from __future__ import annotations

import numpy as np

SYSTEM_PROMPT_OFFSET = 58
TOKEN_IDS = {
    "<|im_start|>": 151644,
    "<|im_end|>": 151645,
    "user": 872,
    "assistant": 77091,
    "<|vision_start|>": 151652,
    "<|vision_end|>": 151653,
    "<|video_pad|>": 151656,
    "\n": 198,
    "previous text": [19702, 1467],
    "Time": 1462,
}

_LABELS = {
    "user": ([TOKEN_IDS["<|im_start|>"], TOKEN_IDS["user"]], TOKEN_IDS["<|im_end|>"]),
    "previous text": ([TOKEN_IDS["<|im_start|>"]] + TOKEN_IDS["previous text"] + [TOKEN_IDS["\n"]], TOKEN_IDS["<|im_end|>"]),
    "user_text": ([TOKEN_IDS["Time"]], TOKEN_IDS["<|vision_start|>"]),
    "assistant": ([TOKEN_IDS["<|im_start|>"], TOKEN_IDS["assistant"]], TOKEN_IDS["<|im_end|>"]),
    "vision": ([TOKEN_IDS["<|vision_start|>"]], TOKEN_IDS["<|vision_end|>"]),
}


def as_host_ids(input_ids) -> np.ndarray:
    """Flat int64 numpy view of a list / numpy array / (CPU or CUDA) torch tensor."""
    if isinstance(input_ids, np.ndarray):
        return input_ids.reshape(-1).astype(np.int64, copy=False)
    if hasattr(input_ids, "detach"):
        return input_ids.detach().reshape(-1).cpu().numpy().astype(np.int64, copy=False)
    return np.asarray(input_ids, dtype=np.int64).reshape(-1)


def find_segments(ids: np.ndarray, label: str, contain_lf: bool = True):
    """All closed (start, end) segments of `label`, left to right, non-overlapping."""
    assert label in _LABELS, label
    pat, end_tok = _LABELS[label]
    n, m = ids.shape[0], len(pat)
    if n < m + 1:
        return []
    hit = ids[: n - m + 1] == pat[0]
    for j in range(1, m):
        hit &= ids[j: n - m + 1 + j] == pat[j]
    starts = np.flatnonzero(hit)
    ends = np.flatnonzero(ids == end_tok)
    if starts.size == 0 or ends.size == 0:
        return []
    lf = TOKEN_IDS["\n"]
    # terminator of every candidate start, all at once; a start without one ends the scan (reference :80-81)
    e_idx = np.searchsorted(ends, starts + m, side="left")
    live = e_idx < ends.shape[0]
    starts, e_idx = starts[live], e_idx[live]
    if starts.size == 0:
        return []
    e = ends[e_idx]
    if starts.size > 1 and not bool(np.all(starts[1:] > e[:-1])):
        # a candidate begins inside the previous segment (adversarial input): fall back to the sequential greedy scan
        keep, cur = [], 0
        for k in range(starts.shape[0]):
            if starts[k] >= cur:
                keep.append(k)
                cur = int(e[k]) + 1
        starts, e = starts[keep], e[keep]
    if contain_lf:
        nxt = np.minimum(e + 1, n - 1)
        e = np.where((e + 1 < n) & (ids[nxt] == lf), e + 1, e)
    return list(zip(starts.tolist(), e.tolist()))


def find_segment(ids: np.ndarray, label: str, index: int, contain_lf: bool = True):
    """The `index`-th segment of `label` (closed interval, before the `user_text` adjustment): one native pass over the ids
    (`svlm_find_segment`, a host function of libsvlm_b200.so -- ~5 us for 5.8k ids where the vectorised numpy scan needs ~100 us, and
    the eviction planner asks five times per chunk).  Same result as `find_segments(ids, label, contain_lf)[index]`, IndexError
    included (`tests/test_host_logic.py` checks the two against each other and against the reference's fixtures)."""
    assert label in _LABELS, label
    from .. import _abi
    pat, end_tok = _LABELS[label]
    if index < -256:                      # beyond the native scan's ring of trailing segments: the vectorised scan answers
        return find_segments(ids, label, contain_lf)[index]
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    return _abi.find_segment(ids, pat, end_tok, index, contain_lf, TOKEN_IDS["\n"])


def get_qwen_range(input_ids, label: str, index: int, contain_lf: bool = True):
    s, e = find_segment(as_host_ids(input_ids), label, index, contain_lf)
    if label == "user_text":
        return s, e - 1
    return s, e
