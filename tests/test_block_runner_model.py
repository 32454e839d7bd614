"""CPU model of the PLL block runner (csrc/pdt_pll_pipe.cuh) — tests/host/block_runner_model.cu replays its schedule (speculated
sweep flag and latch, one block of look-ahead, roll-back to the first contradicted sample, short blocks afterwards) on the host
with the kernels' own per-sample functions and must reproduce the one-thread loop bit for bit: outputs, lock values, phase and
frequency traces, final state — for float and double, sane and absurd loop bandwidths, calls cut at odd lengths."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import pyoracle as po
from tests.synth_ref import make_argos_capture, make_poes_capture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "project-desert-tortoise_b200", "csrc")


@pytest.fixture(scope="module")
def models(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not found: the host model is built from the CUDA headers")
    d = tmp_path_factory.mktemp("brm")
    out = {}
    for prec, fl in (("f32", 1), ("f64", 0)):
        exe = d / f"block_runner_model_{prec}"
        subprocess.run(["nvcc", "-std=c++17", "-O2", "-fmad=false", "-Xcompiler", "-ffp-contract=off", f"-DPDT_USE_FLOATS={fl}",
                        "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "host", "block_runner_model.cu")], check=True)
        out[prec] = exe
    return out


CASES = [  # fs, bw_acq, bw_track, lock_thresh, cuts
    (50000, 0.0126, 0.00126, 0.1, []),                               # near the POES defaults, one call
    (50000, 0.0126, 0.00126, 0.1, [10000] * 5 + [1, 7, 129, 4096]),  # the reference's chunk, then odd lengths
    (50000, 0.3, 0.05, 0.2, [7001, 1, 7998]),                        # outside pll_fast_ok: reference-shaped loops
    (50000, 0.002, 0.0002, 0.05, [3333] * 9),
]


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("fs,bw_acq,bw_track,thresh,cuts", CASES)
def test_block_runner_model_equals_the_one_thread_loop(models, tmp_path, prec, fs, bw_acq, bw_track, thresh, cuts):
    o = po.Oracle(prec)
    pcm, _ = make_poes_capture(60000, fs, 31, esn0_db=13.0, doppler_hz=1500.0, amplitude=0.3)
    iq = o.pcm16_to_complex(pcm)
    iq[: 2 * 9000] *= 0.02                                            # noise-like lead-in: the sweep runs before the signal appears
    raw = tmp_path / "iq.bin"
    np.ascontiguousarray(iq).tofile(raw)
    r = subprocess.run([str(models[prec]), str(fs), str(raw), str(bw_acq), str(bw_track), str(thresh), *map(str, cuts)],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-300:] + r.stderr[-300:]
    f = dict(kv.split("=") for kv in r.stdout.split()[1:])
    assert int(f["samples"]) == 60000 and int(f["blocks"]) >= 60000 // 128
    assert int(f["rolled"]) >= 1                                      # the latch (and usually sweep-flag flips) forced roll-backs


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("thresh,want_locked", [(0.99, 0), (0.35, 1)])
def test_block_runner_model_sweep_flag_contradictions(models, tmp_path, prec, thresh, want_locked):
    """The other kind of contradiction: with the reference's default gains and a latch that fires late (0.35) or never (0.99),
    the loop pulls in while still in acquisition, the averaged phase leaves the noise band and the speculated sweep flag is
    wrong for the first time in the middle of a block — roll-back without a latch (and, for 0.35, the latch later on)."""
    o = po.Oracle(prec)
    pcm, _ = make_poes_capture(100000, 50000, 1, esn0_db=12.0, doppler_hz=-1000.0, amplitude=0.25)
    raw = tmp_path / "iq.bin"
    np.ascontiguousarray(o.pcm16_to_complex(pcm)).tofile(raw)
    r = subprocess.run([str(models[prec]), "50000", str(raw), "0.016", "0.0013", str(thresh), "10000", "10000", "12345"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-300:] + r.stderr[-300:]
    f = dict(kv.split("=") for kv in r.stdout.split()[1:])
    assert int(f["locked"]) == want_locked and int(f["rolled"]) >= 1 + want_locked


def test_block_runner_model_on_an_argos_burst_capture(models, tmp_path):
    o = po.Oracle("f64")
    pcm, _ = make_argos_capture(40000, 5000.0, seed=8, n_bursts=2, snr_db=18.0)
    raw = tmp_path / "iq.bin"
    np.ascontiguousarray(o.pcm16_to_complex(pcm)).tofile(raw)
    r = subprocess.run([str(models["f64"]), "5000", str(raw), "0.0377", "0.00377", "0.1", "2400", "2400", "2400"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-300:] + r.stderr[-300:]
