"""Measured parity numbers (greedy mismatches, maximum errors) of the GPU tests, appended to `gpurun_out/parity_r2.jsonl` so that
the counts the assertions bound are also on record (the driver merges `gpurun_out/` back)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_margins_ulps = []          # oracle margins (in bf16 ulps of the winning logit) of the mismatches seen since the last record()


def record(test: str, **values):
    line = {"test": test, **{k: (float(v) if isinstance(v, float) else v) for k, v in values.items()}}
    if _margins_ulps:
        line["mismatch_margins_bf16_ulps"] = [round(m, 2) for m in _margins_ulps]
        _margins_ulps.clear()
    print("PARITY", json.dumps(line))
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_r2.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    except OSError:
        pass


def near_tie(step_logits, chosen: int, oracle: int) -> bool:
    """A differing greedy token is acceptable only if the ORACLE itself nearly ties the two candidates: its margin must be within two
    bf16 ulps of the winning logit.  The lm_head output is bf16 (8 mantissa bits: the top logits of a random-init model, ~4.5 sigma,
    sit on a 0.03-wide grid while the mean gap between the two best of 152k candidates is ~0.2), so exact ties and one-ulp
    differences between two correct bf16 evaluations are a few per cent of all tokens -- each one is checked, the count is recorded."""
    import math
    top = float(step_logits[oracle])
    ulp = 2.0 ** (math.floor(math.log2(max(abs(top), 1e-30))) - 7)
    margin = float(step_logits[oracle] - step_logits[chosen])
    _margins_ulps.append(margin / ulp)
    return margin <= 2.0 * ulp + 1e-6


MAX_NEAR_TIE_FRACTION = 0.05
