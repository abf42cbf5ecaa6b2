"""Measured parity numbers (greedy mismatches, maximum errors) of the GPU tests, appended to `gpurun_out/parity_r2.jsonl` so that
the counts the assertions bound are also on record (the driver merges `gpurun_out/` back)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(test: str, **values):
    line = {"test": test, **{k: (float(v) if isinstance(v, float) else v) for k, v in values.items()}}
    print("PARITY", json.dumps(line))
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_r2.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    except OSError:
        pass
