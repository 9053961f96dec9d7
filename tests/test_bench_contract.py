"""bench.py's JSON-line contract, checked on the CPU through the reference arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    return out


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run(None, "--impl", "reference", "--steps", "5", "--warmup", "0", "--variant", "tiny")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "30s-clips/sec training" and d["unit"] == "clips/s"
    # the arm runs REAL full-depth steps, at most 3 of them, and reports the count it executed (not the count requested)
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["steps_requested"] == 5 and d["vs_baseline"] is None
    assert "extrapolat" not in d["cpu_baseline"]["sample"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - 1.0) < 1e-6        # one clip per step
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_non_zero_ranks():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
