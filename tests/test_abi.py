"""C-ABI surface checks that need no GPU: the shared objects load, export every symbol declared in
include/*.h, mirror the reference's defaults, and FAIL LOUDLY (no CPU fallback) without a device."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

pdt = importlib.import_module("project-desert-tortoise_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_exports_every_declared_symbol(prec):
    if not os.path.exists(pdt.lib_path(prec)):
        pdt.build()
    lib = C.CDLL(pdt.lib_path(prec))
    declared = _declared("pdt.h") + _declared("pdt_legacy.h")
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by libpdt_{prec}.so"
    assert sorted(set(pdt.EXPORTED_SYMBOLS)) == sorted(set(declared))
    L = pdt.load(prec)
    assert L.pdt_real_size() == (4 if prec == "f32" else 8)
    assert b"sm_100a" in L.pdt_version()


def test_defaults_mirror_reference_drivers():
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, 50000)
    assert (p.chunk, p.interp, p.taps, p.sync_len) == (10000, 3, 78, 19)
    assert p.sync_word == b"1110110111100010000" and abs(p.baud - 16640.3) < 1e-9
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, 250000)
    assert (p.interp, p.taps) == (1, 26)
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, 2000000)
    assert p.interp == 0                                  # the reference's L=0 case
    p = pdt.default_params("f64", pdt.PDT_MODE_ARGOS, 5000)
    assert (p.chunk, p.interp, p.taps, p.sync_len, p.sync_word) == (2400, 1, 50, 13, b"0001011110000")


def test_struct_layouts():
    assert C.sizeof(pdt.Frame) == 120 and pdt.FRAME_DTYPE.itemsize == 120
    assert C.sizeof(pdt.Stats) == pdt.STATS_DTYPE.itemsize
    for name, _ in pdt.Stats._fields_:
        assert getattr(pdt.Stats, name).offset == pdt.STATS_DTYPE.fields[name][1]
    for name, _ in pdt.Frame._fields_:
        assert getattr(pdt.Frame, name).offset == pdt.FRAME_DTYPE.fields[name][1]


def test_scalar_helpers_match_oracle(oracle32):
    L = pdt.load("f32")
    rng = np.random.default_rng(3)
    for y, x in rng.standard_normal((500, 2)):
        assert L.arctan2(y, x) == oracle32.lib.pdto_arctan2(y, x)
    for x in np.abs(rng.standard_normal(500)) * 100:
        assert L.Q_rsqrt(x) == oracle32.lib.pdto_q_rsqrt(x)
    assert [L.sign(v) for v in (-2.0, 0.0, 3.0)] == [-1, 0, 1]
    h = pdt.Legacy("f32").MakeLPFIR(78, 11000.0, np.float32(150000.0), 3)
    assert np.array_equal(h, oracle32.make_lpfir(78, 11000.0, np.float32(150000.0), 3))


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product refuses to run — it never falls back to a CPU path."""
    L = pdt.load("f32")
    if L.pdt_device_count() > 0:
        pytest.skip("a GPU is present")
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, 250000)
    with pytest.raises(pdt.PdtError, match="CUDA"):
        pdt.Demod("f32", p, 1, 1000)
    assert L.pdt_set_device(0) != 0
