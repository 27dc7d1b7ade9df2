"""The JSON line of bench.py: the reference arm runs here (CPU), the B200 arm is checked on its last committed line
(profiles/), so that a key of the contract cannot go missing unnoticed."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-frames", "1",
                        "--width", "256", "--height", "160"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["impl"] == "reference" and line["metric"] == "frames/sec" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert "workload" in line["config"] and "model" not in line["config"]


import pytest


@pytest.mark.parametrize("name", ["r01_v8_bench.json", "r02_z_bench.json"])
def test_committed_b200_line_has_the_contract_keys(name):
    line = json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(line)
    assert line["n_gpus"] == 1 and line["data"] == "synthetic" and line["dtype"] == "f32" and line["vs_baseline"] is None
    if name.startswith("r02"):  # the round-2 line: roofline of the dominant kernel, kernels inside the timed region
        r = line["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
        assert line["gpu_launches"] == 3 * 60 * line["steps"] and line["e2e"]["h2d_bytes_per_step"] > 0
    rf = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "hbm" and rf["unit"] == "GB/s"
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] == 60 * 4 * 1920 * 1080 * 12 and e["d2h_bytes_per_step"] == 60 * 1920 * 1080 * 12
    assert 0 < e["value"] < line["value"]
    assert line["gpu_launches"] == 3 * 60 * line["steps"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
