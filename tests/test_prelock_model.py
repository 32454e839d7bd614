"""CPU model of the principle behind the tiled engine's PLL tiles and the stream mode's pre-locked segments:
in TRACK mode the reference's loop (CarrierTrackingPLL.c:165-188 with the gains of :272-273) forgets its start state — a
trajectory started from a *guess* becomes BIT-identical to the serial one within the warm-up window W = 17 / bw_track the
kernels use, and stays identical.  Checked here with the CPU oracle only (no GPU): this is the property, the GPU tests check
the kernels."""
import ctypes as C

import numpy as np
import pytest

from tests.synth_ref import make_poes_capture


class PllS(C.Structure):            # oracle/pdt_oracle.h pdto_pll (float build)
    _fields_ = [("first_lock", C.c_long), ("damp", C.c_float), ("alpha", C.c_float), ("beta", C.c_float),
                ("phase", C.c_float), ("freq", C.c_float), ("max_freq", C.c_float), ("min_freq", C.c_float),
                ("avg_phase", C.c_float), ("locksig", C.c_float), ("sweep", C.c_float),
                ("lock_freq_hz", C.c_double), ("samples_seen", C.c_uint64), ("lock_sample", C.c_uint64)]


def _run(oracle, st, iq, fs, start, stop, chunk=10000):
    w = 2.0 * np.pi / fs
    ph, fr = [], []
    for pos in range(start, stop, chunk):
        m = min(chunk, stop - pos)
        _, _, _, tp, tf = oracle.pll(st, iq[2 * pos: 2 * (pos + m)], fs, 4500.0, 0.08, 0.3979 * w, 127.3240 * w, 10.3451 * w, trace=True)
        ph.append(tp)
        fr.append(tf)
    return np.concatenate(ph), np.concatenate(fr)


@pytest.mark.parametrize("fs,d_phase,d_hz", [(250000, 0.2, 2.0), (250000, -1.0, -10.0), (250000, 2.5, 25.0), (50000, 0.2, 2.0)])
def test_track_mode_pll_forgets_its_start_state_bit_exactly(oracle32, fs, d_phase, d_hz):
    n = int(2.4 * fs)
    pcm, _ = make_poes_capture(n, fs, 5, esn0_db=14.0, doppler_hz=700.0, drift_hz_s=30.0, amplitude=0.2)
    iq = oracle32.pcm16_to_complex(pcm)
    assert oracle32.lib.pdto_sizeof(b"pll") == C.sizeof(PllS)
    k0 = (n // 2) // 10000 * 10000                           # a chunk boundary well behind the lock latch
    st = oracle32.new_state("pll")
    _run(oracle32, st, iq, fs, 0, k0)
    s = C.cast(st.p, C.POINTER(PllS)).contents
    assert s.first_lock >= 0 and s.lock_sample < k0            # latched: the loop is in track mode from here on
    snap = C.string_at(st.p, C.sizeof(PllS))
    ph_serial, fr_serial = _run(oracle32, st, iq, fs, k0, n)
    # the same state with the phase / frequency a carrier estimator could be off by
    st2 = oracle32.new_state("pll")
    C.memmove(st2.p, snap, C.sizeof(PllS))
    s2 = C.cast(st2.p, C.POINTER(PllS)).contents
    s2.phase = np.float32(s2.phase + d_phase)
    s2.freq = np.float32(s2.freq + 2.0 * np.pi * d_hz / fs)
    ph_guess, fr_guess = _run(oracle32, st2, iq, fs, k0, n)
    differ = np.nonzero((ph_guess != ph_serial) | (fr_guess != fr_serial))[0]
    assert differ.size > 0 and differ[0] == 0                  # it really started somewhere else
    merged_at = int(differ[-1]) + 1                            # first sample from which both trajectories are the same bits
    W = int(17.0 / (10.3451 * 2.0 * np.pi / fs))
    print("merged after", merged_at, "samples; warm-up window", W)
    assert merged_at <= W
    assert ph_serial.size - merged_at > 2 * W                  # and stayed merged for a long stretch behind it


class AgcS(C.Structure):            # oracle/pdt_oracle.h pdto_agc
    _fields_ = [("init", C.c_int), ("gain", C.c_float)]


@pytest.mark.parametrize("fs,scale", [(250000, 1.3), (250000, 0.7), (50000, 1.2)])
def test_agc_forgets_its_start_gain_bit_exactly(oracle32, fs, scale):
    """Same property for NormalizingAGC (AGC.c:98-131): started from a gain that is 20-30 % off — what 1/mean|y| over a few
    hundred samples gives (k_agc_plan) — the gain stream becomes bit-identical to the serial one within the kernels' warm-up
    W = 22·gain/decay interpolated samples."""
    n = int(2.4 * fs)
    pcm, _ = make_poes_capture(n, fs, 6, esn0_db=14.0, doppler_hz=-900.0, amplitude=0.2)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, trace=True)
    y = want["tr_lpf"]                                          # FIR output of the serial chain = the AGC's input
    L = int(np.rint(150000.0 / fs))
    assert oracle32.lib.pdto_sizeof(b"agc") == C.sizeof(AgcS)
    norm = oracle32.static_gain(iq[: 2 * 10000])
    w = 2.0 * np.pi / float(np.float32(fs) * np.float32(L))      # double division of the float product, like the chain
    attack, decay = 79.5775 * w, 159.1549 * w                   # main.c:97-98 as rad/s over the interpolated rate
    chunk = 10000 * L
    k0 = (y.size // 2) // chunk * chunk
    st = oracle32.new_state("agc")
    for pos in range(0, k0, chunk):
        oracle32.agc(st, y[pos: pos + chunk], norm, attack, decay)
    snap = C.string_at(st.p, C.sizeof(AgcS))

    def tail(state):
        outs = []
        for pos in range(k0, y.size, chunk):
            z, _ = oracle32.agc(state, y[pos: pos + chunk], norm, attack, decay)
            outs.append(z)
        return np.concatenate(outs)

    z_serial = tail(st)
    assert np.array_equal(z_serial, want["tr_agc"][k0:])        # the stage-wise replay is the chain's AGC stream
    st2 = oracle32.new_state("agc")
    C.memmove(st2.p, snap, C.sizeof(AgcS))
    s2 = C.cast(st2.p, C.POINTER(AgcS)).contents
    g0 = float(s2.gain)
    s2.gain = np.float32(g0 * scale)
    z_guess = tail(st2)
    differ = np.nonzero(z_guess != z_serial)[0]
    assert differ.size > 0 and differ[0] == 0
    merged_at = int(differ[-1]) + 1
    W = int(22.0 * max(g0, 0.25) / decay)
    print("AGC merged after", merged_at, "samples; warm-up window", W, "gain", g0)
    assert merged_at <= W and z_serial.size - merged_at > W
