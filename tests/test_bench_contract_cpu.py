"""The driver's contract for `bench.py`, checked without a GPU: the module imports on a CPU-only box, the defaults are the ones the
driver relies on (N = 1, >= 3 warm-up steps), and the bench line committed at the end of the round carries every key the contract
names, with self-consistent roofline arithmetic."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_module_imports_and_defaults():
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    assert bench.UNIT == "frames/s" and bench.NEW_TOKENS == 20
    src = open(os.path.join(ROOT, "bench.py")).read()
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in src
    assert 'ap.add_argument("--gpus", type=int, default=1)' in src


@pytest.mark.parametrize("name", ["r2l_bench_line_final.json", "r2l_scale_8gpu.json"])
def test_committed_bench_line_carries_the_contract(name):
    path = os.path.join(ROOT, "profiles", name)
    line = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == "steady_state_frames_per_s" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 1_000_000 and e2e["d2h_bytes_per_step"] > 0 and e2e["unit"] == "frames/s"
    assert abs(e2e["value"] - line["value"]) / line["value"] < 0.05 and e2e["value"] != line["value"]
    clocks = line["clocks"]
    assert clocks["sm_mhz"] <= clocks["sm_max_mhz"] and not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if line["n_gpus"] == 1:
        roof = line["roofline"]
        assert roof["bound"] == "hbm" and roof["unit"] == "GB/s"
        assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and 0.5 < roof["frac"] < 1.2
        assert roof["traffic"] is None or roof["traffic"] >= 0.95 * 271_581_184        # DRAM bytes >= the algorithmic bytes of the launch
        cpu = line["cpu_baseline"]
        assert cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1 and cpu["unit"] == "frames/s" and cpu["sample"]
        for leg in ("multi_stream", "multi_stream_16"):
            assert line[leg]["value"] > line["value"] and "decode_projections" in line[leg]
    else:
        assert line["value"] > 7 * 30.0                       # eight independent streams, no collective on the data path
