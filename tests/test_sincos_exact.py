"""The kernels' sinf/cosf (pdt_device.cuh::sincos_core) restated in numpy, op for op, against glibc's sinf/cosf.

The PLL derotates every sample with sinf/cosf of a phase in [-2π, 2π] (CarrierTrackingPLL.c:106-107); bit-exact PLL output
needs the device to reproduce glibc's result for every such float.  sincos_core evaluates glibc's own algorithm (ARM
optimized-routines: double polynomial after a 2^24·2/π quadrant reduction, one rounding to float) as ONE branch-free path:
no small-argument shortcut, signs applied to the float results with XOR masks from the quadrant bits.  Checked here on the
CPU (IEEE double arithmetic is the same on both sides): every float of [2^-13, 8] of both signs, the region around glibc's
2^-12 shortcut, samples of the tiny range, and the special values; y = -0 is the documented exception the kernels exclude.
"""
import ctypes as C
import os

import numpy as np

import pyoracle as po

S = [float.fromhex(h) for h in ("-0x1.555545995a603p-3", "0x1.1107605230bc4p-7", "-0x1.994eb3774cf24p-13")]
Cc = [float.fromhex(h) for h in ("0x1p0", "-0x1.ffffffd0c621cp-2", "0x1.55553e1068f19p-5", "-0x1.6c087e89a359dp-10", "0x1.99343027bf8c3p-16")]
HPI_INV, HPI = float.fromhex("0x1.45F306DC9C883p+23"), float.fromhex("0x1.921FB54442D18p0")


def sincos_core_np(y):
    """pdt_device.cuh::sincos_core, one numpy op per device op (float64 = IEEE double, float32 conversions round to nearest)."""
    x = y.astype(np.float64)
    r = x * HPI_INV
    t1 = np.trunc(r).astype(np.int64).astype(np.int32) + np.int32(0x800000)
    n = t1 >> 24
    xr = x - n.astype(np.float64) * HPI
    x2 = xr * xr
    x3 = xr * x2
    t1s = S[1] + x2 * S[2]
    x7 = x3 * x2
    sp = (xr + x3 * S[0]) + x7 * t1s
    x4 = x2 * x2
    t2 = Cc[3] + x2 * Cc[4]
    t1c = Cc[0] + x2 * Cc[1]
    x6 = x4 * x2
    cp = (t1c + x4 * Cc[2]) + x6 * t2
    tu = t1.view(np.uint32)
    ms = ((tu + np.uint32(0x1000000)) << np.uint32(6)) & np.uint32(0x80000000)
    mc = (tu << np.uint32(6)) & np.uint32(0x80000000)
    sv = (sp.astype(np.float32).view(np.uint32) ^ ms).view(np.float32)
    cv = (cp.astype(np.float32).view(np.uint32) ^ mc).view(np.float32)
    odd = (t1 & np.int32(0x1000000)) != 0
    return np.where(odd, cv, sv), np.where(odd, sv, cv)


def glibc(y):
    lib = po.Oracle("f32").lib
    lib.pdto_sincosf_array.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    y = np.ascontiguousarray(y, np.float32)
    s, c = np.empty_like(y), np.empty_like(y)
    lib.pdto_sincosf_array(y.ctypes.data, y.size, s.ctypes.data, c.ctypes.data)
    return s, c


def _check(y):
    s, c = sincos_core_np(y)
    gs, gc = glibc(y)
    bad = np.nonzero((s.view(np.uint32) != gs.view(np.uint32)) | (c.view(np.uint32) != gc.view(np.uint32)))[0]
    assert bad.size == 0, (bad.size, y[bad[:5]], s[bad[:5]], gs[bad[:5]], c[bad[:5]], gc[bad[:5]])


def test_every_float_from_2pow_minus13_to_8_both_signs():
    lo, hi = np.float32(2.0 ** -13).view(np.uint32), np.float32(8.0).view(np.uint32)
    step = 1 << 23
    for a in range(int(lo), int(hi), step):                               # one binade (8.4 M floats) at a time
        y = np.arange(a, min(a + step, int(hi) + 1), dtype=np.uint32).view(np.float32)
        _check(y)
        _check(-y)


def test_tiny_arguments_and_special_values():
    rng = np.random.default_rng(3)
    bits = rng.integers(1, int(np.float32(2.0 ** -13).view(np.uint32)), 6_000_000, dtype=np.uint32)   # denormals … 2^-13
    y = bits.view(np.float32)
    _check(y)
    _check(-y)
    edge = np.float32([0.0, 1e-45, -1e-45, 2.0 ** -126, 2.0 ** -12, np.nextafter(np.float32(2.0 ** -12), np.float32(0)),
                       np.pi / 4, np.pi / 2, np.pi, 2 * np.pi, -2 * np.pi, 6.2831855, 100.0, -119.99, 119.99])
    _check(edge)
    # the documented exception: the sine kernel returns +0 for y = -0 where glibc returns -0 (sincos_in_core_range excludes it)
    s, _ = sincos_core_np(np.float32([-0.0]))
    gs, _ = glibc(np.float32([-0.0]))
    assert s.view(np.uint32)[0] == 0 and gs.view(np.uint32)[0] == 0x80000000


def test_the_kernels_own_sincos_on_the_host_equals_glibc_for_every_phase(tmp_path):
    """Not a restatement: csrc/pdt_device.cuh::sincos_core itself (`__host__ __device__`), compiled by nvcc into a CPU program
    (tests/host/sincos_host.cu) and compared with glibc's sinf / cosf for every float of ±[2^-13, 8] plus samples of the rest of
    its declared range — 270 M values."""
    import os
    import shutil
    import subprocess
    import pytest
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not found: the host harness is built from the CUDA headers")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "sincos_host"
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-fmad=false", "-Xcompiler", "-ffp-contract=off", "-DPDT_USE_FLOATS=1",
                    "-I" + os.path.join(root, "project-desert-tortoise_b200", "csrc"), "-o", str(exe),
                    os.path.join(root, "tests", "host", "sincos_host.cu")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-300:]
    assert int(r.stdout.split()[1]) > 250_000_000

