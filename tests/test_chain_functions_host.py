"""The per-sample functions of csrc/pdt_device.cuh (`__host__ __device__`: pll_step, FIR, agc_step, gardner_step, manchester_step,
sync_step, StaticGain) compiled by nvcc into a CPU program (tests/host/chain_host.cu) that calls them in the reference's order,
against the oracle: the golden 50 ksps clip (float, L = 3) and a synthetic ARGOS capture (double).  No device: this keeps the
arithmetic the kernels are built from under the CPU suite; the kernels themselves are the GPU tests' business."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import pyoracle as po
from tests.synth_ref import make_argos_capture, parse_frames_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "project-desert-tortoise_b200")
CSRC = os.path.join(PKG, "csrc")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _build(tmp_path, prec):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not found: the host harness is built from the CUDA headers")
    if not os.path.exists(os.path.join(PKG, f"libpdt_{prec}.so")):
        pytest.skip(f"libpdt_{prec}.so not built")
    exe = tmp_path / f"chain_host_{prec}"
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
                    f"-DPDT_USE_FLOATS={1 if prec == 'f32' else 0}", "-I" + CSRC, "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "chain_host.cu"), "-L" + PKG, f"-lpdt_{prec}",
                    "-Xlinker", "-rpath," + PKG], check=True)
    return exe


def _run(exe, mode, fs, iq, tmp_path):
    raw = tmp_path / "iq.bin"
    np.ascontiguousarray(iq).tofile(raw)
    r = subprocess.run([str(exe), str(mode), str(fs), str(raw)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    frames, totals = [], None
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "F":
            frames.append((bool(int(t[1])), bytes.fromhex(t[3]) if len(t) > 3 else b""))
        elif t[0] == "T":
            totals = tuple(int(v) for v in t[1:])
    return frames, totals


def test_poes_clip_through_the_device_functions_on_the_host(tmp_path):
    exe = _build(tmp_path, "f32")
    rate, pcm = po.read_wav_pcm16(os.path.join(GOLDEN, "5sec_clip.wav"))
    o = po.Oracle("f32")
    iq = o.pcm16_to_complex(pcm)
    want = o.chain(iq, rate)
    frames, totals = _run(exe, 0, rate, iq, tmp_path)
    assert totals[:3] == (want["total_symbols"], want["total_bits"], want["total_frames"]) and totals[3] == 1
    rows = [(r[1], bytes(r[2])) for r in parse_frames_text(want["text"])]
    assert len(rows) > 40 and [f for f in frames] == rows
    golden = [(r[1], bytes(r[2])) for r in parse_frames_text(open(os.path.join(GOLDEN, "poes_5sec_clip_frames.txt")).read())]
    assert frames == golden                                   # …which is what the unmodified reference wrote for this file


def test_argos_capture_through_the_device_functions_on_the_host(tmp_path):
    exe = _build(tmp_path, "f64")
    o = po.Oracle("f64")
    pcm, _ = make_argos_capture(60000, 5000.0, seed=5, n_bursts=4, snr_db=18.0)
    iq = o.pcm16_to_complex(pcm)
    want = o.chain(iq, 5000, argos=True)
    frames, totals = _run(exe, 1, 5000, iq, tmp_path)
    assert totals[:3] == (want["total_symbols"], want["total_bits"], want["total_frames"])
    rows = [bytes(r[2]) for r in parse_frames_text(want["text"])]
    assert want["total_frames"] >= 2 and [f[1] for f in frames] == rows
