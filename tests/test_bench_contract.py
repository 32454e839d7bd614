"""The JSON lines bench.py prints are a contract with the driver.  bench.py --impl reference needs no GPU: it times the reference's own CPU path (oracle/_ref when the reference compiled here,
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


@pytest.mark.gpu
def test_our_arm_line_on_a_small_batch():
    """The product arm's JSON line on a small batch (64 captures x 300 k samples): every key of the contract, the roofline and
    cpu_baseline objects, a positive launch count, the frame-bytes check of the sampled captures against the reference."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--captures", "64", "--samples", "300000", "--steps", "4", "--warmup", "3",
                        "--no-single", "--cpu-captures", "4"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    need = (REQUIRED - {"impl"}) | {"gpu_launches", "roofline", "clocks"}
    assert need <= set(b), need - set(b)
    assert b["value"] > 0 and b["n_gpus"] == 1 and b["gpu_launches"] > 0 and b["dtype"] == "f32" and b["scaling"] == "weak"
    rf = b["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "hbm" and 0 < rf["frac"] < 1.2
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = b["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 64 * 300000 * 8 and e["d2h_bytes_per_step"] > 0
    cb = b["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1
    assert cb.get("frame_bytes_check", {}).get("equal") is True
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(b["clocks"])
