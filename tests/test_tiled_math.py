"""Exhaustive CPU checks of the float-only identities the tiled engine relies on (pdt_tiled.cuh):

  * (float)((double)d - 2*M_PI) == (d - HI) - MID   for EVERY float d in [3, 10.5]   (and mirrored for +2π)
  * (double)d >  M_PI    <=>  d >= PI_UP
  * (double)p >  2*M_PI  <=>  p >= TWO_PI_HI
so that the reference's double-precision wraps (CarrierTrackingPLL.c:165-182) can run as two float adds and a select.
Also the integer "rint" used by the Gardner kernels and the round-to-float-in-double used in studies.
"""
import numpy as np

HI = np.float32(float.fromhex("0x1.921fb6p+2"))           # float(2π)
MID = np.float32(float.fromhex("-0x1.777a5cp-23"))       # float(2π - HI)
PI_UP = np.float32(float.fromhex("0x1.921fb6p+1"))        # float(π) — the first float above π
TWO_PI = 2.0 * np.pi


def _all_floats(lo, hi):
    a, b = np.float32(lo).view(np.uint32), np.float32(hi).view(np.uint32)
    return np.arange(int(a), int(b) + 1, dtype=np.uint32).view(np.float32)


def test_wrap_down_up_exact_for_every_float_in_range():
    d = _all_floats(3.0, 10.5)
    assert d.size > 15_000_000
    want = (d.astype(np.float64) - TWO_PI).astype(np.float32)               # what the reference computes
    got = (d - HI) - MID                                                    # float32 arithmetic, op by op
    assert got.dtype == np.float32 and np.array_equal(want, got)
    # mirrored: (float)((double)d + 2π) for d in [-10.5, -3]
    dn = -d
    want_up = (dn.astype(np.float64) + TWO_PI).astype(np.float32)
    got_up = (dn + HI) + MID
    assert np.array_equal(want_up, got_up)


def test_threshold_constants_are_the_first_floats_above_pi_and_two_pi():
    assert float(PI_UP) > np.pi and float(np.nextafter(PI_UP, np.float32(0))) < np.pi
    assert float(HI) > TWO_PI and float(np.nextafter(HI, np.float32(0))) < TWO_PI
    # hence for floats: (double)d > M_PI <=> d >= PI_UP, (double)p > 2*M_PI <=> p >= HI
    d = _all_floats(3.0, 3.3)
    assert np.array_equal(d.astype(np.float64) > np.pi, d >= PI_UP)
    p = _all_floats(6.0, 6.6)
    assert np.array_equal(p.astype(np.float64) > TWO_PI, p >= HI)
    assert float(MID) == float(np.float32(TWO_PI - float(HI)))


def test_magic_number_rint_equals_rintf():
    """k_gardner: rintf(x) as an integer = bits(x + 1.5·2^23) - 0x4B400000 for |x| < 2^22 (ties to even, like rintf)."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-0.5, 40000.0, 2_000_000).astype(np.float32),
                        (np.arange(0, 200000, dtype=np.float32) * np.float32(0.5)),          # exact ties
                        _all_floats(9999.0, 10001.0), np.float32([0.0, -0.0, -0.25, -0.5, 0.5, 1.5, 2.5, 4194303.0])])
    magic = (x + np.float32(12582912.0)).view(np.int32) - np.int32(0x4B400000)
    assert np.array_equal(magic, np.rint(x).astype(np.int32))


def test_float_sum_equals_double_sum_for_gardner_midpoint():
    """k_gardner: next + step/2 computed in float equals the reference's double sum narrowed to float
    (GardenerClockRecovery.c:59) whenever |next| >= 2^-20 or next == 0 (the kernel re-runs the batch otherwise)."""
    rng = np.random.default_rng(2)
    step = np.float32(np.float32(250000) / np.float32(16640.3))
    half = np.float32(np.float64(step) / 2.0)
    assert np.float64(half) == np.float64(step) / 2.0
    nxt = np.concatenate([rng.uniform(-1.0, 31000.0, 3_000_000).astype(np.float32),
                          (rng.uniform(-1, 1, 500_000) * 1e-3).astype(np.float32), np.float32([0.0, 2.0 ** -20, -(2.0 ** -20)])])
    nxt = nxt[(np.abs(nxt) >= np.float32(2.0 ** -20)) | (nxt == 0)]
    want = (nxt.astype(np.float64) + np.float64(step) / 2.0).astype(np.float32)
    assert np.array_equal(want, nxt + half)


def test_sweep_decision_is_an_interval_of_floats():
    """k_acquire_packed: CarrierTrackingPLL.c:232 — fabsf((float)(M_PI/2 - averagePhase)) < 0.05 with averagePhase a float,
    the difference taken in double and narrowed — holds exactly for the floats of ONE interval [lo, hi] (the narrowing is
    monotonic), so the kernel finds lo / hi once by bisection and decides with two float compares.  Checked for every float
    of [1.0, 2.2] and a spread of others."""
    def noise_like(avg):
        d = (np.float64(np.pi / 2.0) - avg.astype(np.float64)).astype(np.float32)
        return np.abs(d).astype(np.float64) < 0.05
    x = _all_floats(1.0, 2.2)
    t = noise_like(x)
    idx = np.nonzero(t)[0]
    assert idx.size > 800_000 and np.array_equal(idx, np.arange(idx[0], idx[-1] + 1))       # one contiguous run of bit patterns
    lo, hi = x[idx[0]], x[idx[-1]]
    assert np.array_equal(t, (x >= lo) & (x <= hi))
    rng = np.random.default_rng(4)
    y = np.concatenate([rng.uniform(-10, 10, 1_000_000).astype(np.float32), np.float32([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, 3e38])])
    with np.errstate(invalid="ignore"):
        assert np.array_equal(noise_like(y), (y >= lo) & (y <= hi))
