"""bench.py's reference arm (`--impl reference`) on CPU: it times the reference's own implementation of the path
(oracle/_ref/libfs2d_ref.so = the unmodified reference sources, Release flags) on a bounded sample and prints one JSON
line with the contract's keys. The GPU arm needs a device and is exercised on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfs2d_ref.so")):
        pytest.skip("oracle/_ref not built")
    # the default resolution (4096^2) costs the reference a minute per substep; the contract is the same at 256^2
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--res", "256", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, FS2D_REF_SECONDARY="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "substeps_per_s" and d["unit"] == "substeps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3 and d["n_gpus"] == 1
    assert abs(d["ms_per_step"] * d["steps"] * 1e-3 * d["value"] - d["steps"]) < 1e-6   # the line describes what was really timed
    assert d["scaling"] == "strong" and d["config"]["cells"] == 256 * 256 and d["config"]["particles"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "256x256" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and not r.stdout.strip()
