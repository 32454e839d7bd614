"""bench.py --impl reference needs no GPU: it times the reference's own CPU path (oracle/_ref when the reference compiled here,
else the oracle port).  The JSON line it prints is a contract with the driver — checked here on a tiny sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _line(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_line_poes():
    b = _line("--ref-captures", "2", "--ref-samples", "200000")
    assert REQUIRED <= set(b) and b["impl"] == "reference" and b["unit"] == "Msamples/s" and b["higher_is_better"] is True
    assert b["value"] > 0 and b["vs_baseline"] is None and b["dtype"] == "f32" and b["data"] == "synthetic"
    assert b["cpu_baseline"]["kind"] in ("reference", "port") and b["cpu_baseline"]["cores"] >= 1 and b["cpu_baseline"]["value"] == b["value"]
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in b["config"] and not any(k in b["config"] for k in ("model", "seq_len", "global_batch"))


def test_reference_arm_line_argos():
    b = _line("--mode", "argos")
    assert REQUIRED <= set(b) and b["impl"] == "reference" and b["dtype"] == "f64" and b["value"] > 0
    assert b["cpu_baseline"]["kind"] in ("reference", "port") and b["e2e"]["h2d_bytes_per_step"] == 0


def test_other_ranks_of_a_reference_run_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=120, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
