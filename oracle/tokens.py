"""Oracle: token-pattern range finder (test infrastructure, CPU only).

Restates `get_qwen_range` of the reference (`utils/get_qwen_range.py:15-86`) and its
token table (`:2-13`).  Closed intervals, greedy left-to-right non-overlapping matching,
optional absorption of a trailing "\\n", negative `index` allowed, `user_text` ends one
token before `<|vision_start|>` (`:84-85`).
"""
from __future__ import annotations

IM_START = 151644
IM_END = 151645
USER = 872
ASSISTANT = 77091
VISION_START = 151652
VISION_END = 151653
VIDEO_PAD = 151656
LF = 198
PREVIOUS_TEXT = (19702, 1467)
TIME = 1462

_PATTERNS = {
    # label: (start pattern, end token)       reference lines
    "user": ((IM_START, USER), IM_END),                        # :38-43
    "previous text": ((IM_START,) + PREVIOUS_TEXT + (LF,), IM_END),  # :44-46
    "user_text": ((TIME,), VISION_START),                      # :47-49
    "assistant": ((IM_START, ASSISTANT), IM_END),              # :50-55
    "vision": ((VISION_START,), VISION_END),                   # :56-58
}


def all_ranges(ids, label: str, contain_lf: bool = True):
    """Every (start, end) segment of `label`, in order (reference :60-82)."""
    if label not in _PATTERNS:
        raise AssertionError(label)
    ids = [int(t) for t in (ids.flatten().tolist() if hasattr(ids, "flatten") else ids)]
    pat, end_tok = _PATTERNS[label]
    n, m = len(ids), len(pat)
    out = []
    pos = 0
    while pos + m <= n:
        if tuple(ids[pos:pos + m]) != pat:
            pos += 1
            continue
        # first end token at or after pos+m; none -> the scan stops altogether (:80-81)
        try:
            e = ids.index(end_tok, pos + m)
        except ValueError:
            break
        if contain_lf and e + 1 < n and ids[e + 1] == LF:
            out.append((pos, e + 1))
        else:
            out.append((pos, e))
        pos = e + 1
    return out


def get_qwen_range(ids, label: str, index: int, contain_lf: bool = True):
    segs = all_ranges(ids, label, contain_lf)
    s, e = segs[index]
    if label == "user_text":
        return s, e - 1
    return s, e
