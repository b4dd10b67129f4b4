"""bench.py on the GPU with its quick settings: one JSON line carrying every key the driver reads, every aux leg present
and none of them reporting an error (a leg that fails is replaced by {"error": ...} so that the headline survives)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_line_has_every_key_and_no_failed_leg(gpu):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--quick-aux", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "aux"):
        assert k in d, k
    assert d["value"] > 1e7 and d["e2e"]["value"] > 1e7 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] > 7e6 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "fp32" and 0.3 < rf["frac"] < 1.0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    legs = ("c3_grid_build", "c3_lookup", "c4_mc", "c5_screen", "single_pose_calls", "c2_fp64_scan", "c2_grid_scan")
    for leg in legs:
        assert leg in d["aux"], leg
        assert "error" not in d["aux"][leg], (leg, d["aux"][leg])
    assert d["aux"]["c5_screen"]["topk_merged"] == 100
    assert d["aux"]["c3_lookup"]["roofline"]["frac"] > 0.3
