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


def _carrier_guess(iq_c, fs, at):
    """What k_estimate / k_prelock compute (pdt_tiled_kernels.cuh est_carrier): decimate the 1024·D samples in front of `at`
    by block sums, 1024-point DFT, 3-bin peak interpolation, then the phase at `at` from a coherent sum of the last 1024
    samples.  Returns (phase, freq in rad/sample)."""
    D = max(int(fs / (2.5 * (4500.0 + 600.0))), 1)
    seg = iq_c[at - 1024 * D: at]
    z = np.fft.fft(seg.reshape(1024, D).sum(axis=1))
    bin_hz = fs / D / 1024.0
    kmax = min(int(5100.0 / bin_hz) + 1, 510)
    ks = np.arange(-kmax, kmax + 1)
    k = int(ks[np.argmax(np.abs(z[ks]) ** 2)])
    xm, x0, xp = z[k - 1], z[k], z[(k + 1) % 1024]
    d = np.real((xm - xp) / (2.0 * x0 - xm - xp))
    f_hz = (k + float(np.clip(d, -0.5, 0.5))) * bin_hz
    cyc = f_hz / fs
    j = np.arange(-1024, 0)
    acc = np.sum(iq_c[at - 1024: at] * np.exp(-2j * np.pi * cyc * j))
    return float(np.angle(acc)), float(2.0 * np.pi * cyc)


def _parse(text, t_offset):
    """frames of the chain's text output -> [(time of the sync bit in stream seconds, bytes)], complete frames only"""
    out = []
    for line in text.splitlines():
        f = line.split()
        if len(f) == 105:
            out.append((float(f[0].rstrip("i")) + t_offset, bytes(int(b, 16) for b in f[1:])))
    return out


def _time_axis(fs, n):
    """wave.c:91-167 accumulates `time += Ts` in float once per sample; the column printed with a frame is that value.
    Returns the accumulated axis so that a printed time can be mapped back to the sample it belongs to."""
    return np.cumsum(np.full(n + 2, np.float32(1.0) / np.float32(fs), np.float32), dtype=np.float32)


def _to_samples(frames, axis, first):
    return [(first + int(np.searchsorted(axis, np.float32(t - 2e-5))), b) for t, b in frames]


@pytest.mark.parametrize("fs,total,segment,esn0,amp,force_l1,min_frames",
                         [(250000, 3_000_000, 500_000, 14.0, 0.2, False, 115),
                          (2_000_000, 12_000_000, 2_000_000, 24.0, 0.3, True, 55)])
def test_stream_as_prelocked_segments_equals_serial_chain_cpu_model(oracle32, fs, total, segment, esn0, amp, force_l1, min_frames):
    """The stream mode of DESIGN §8 restated with the CPU oracle only: a stream cut into segments (+0.3 s lead, +0.13 s
    tail), every segment behind the first started in TRACK mode from the FFT carrier guess, frames kept by ownership
    windows — the stitched list must be the serial chain's list of minor frames, byte for byte.  Second case: BASELINE
    configs[4]'s 2 Msps rate (declared L = max(1, …) deviation, pinned to the patched reference in test_oracle_vs_ref.py) on
    a recording the serial chain locks on."""
    lead, tail = int(0.3 * fs), int(0.13 * fs) + 4096
    pcm, _ = make_poes_capture(total, fs, 9, esn0_db=esn0, doppler_hz=1200.0, drift_hz_s=-150.0, amplitude=amp)
    iq = oracle32.pcm16_to_complex(pcm)
    iq_c = iq.astype(np.float64).view(np.complex128)
    axis = _time_axis(fs, total)
    res = oracle32.chain(iq, fs, force_min_L1=force_l1)
    assert res["locked"]
    serial = _to_samples(_parse(res["text"], 0.0), axis, 0)
    assert len(serial) >= min_frames
    CH = oracle32._chain_struct()
    D = max(int(fs / (2.5 * 5100.0)), 1)
    pre = 1024 * D
    n_seg = -(-(total - lead) // segment)
    stitched = []
    for s in range(n_seg):
        start = s * segment
        stop = min(start + lead + segment + tail, total)
        lo = 0 if s == 0 else start + lead
        hi = total + 1 if s == n_seg - 1 else start + segment + lead
        c = oracle32.lib.pdto_chain_new(0, float(fs), 10000, int(force_l1))
        try:
            first = start
            if s > 0:
                ch = CH.from_address(c)
                phase, freq = _carrier_guess(iq_c, fs, start + pre)
                bw, damp = np.float32(10.3451 * (2.0 * np.pi / fs)), np.float32(0.999)
                p = ch.pll
                p.first_lock, p.damp = 0, damp                                                  # latched: track mode
                p.alpha = (4.0 * damp * bw) / (1.0 + 2.0 * damp * bw + bw * bw)                # CarrierTrackingPLL.c:272-273
                p.beta = (4.0 * bw * bw) / (1.0 + 2.0 * damp * bw + bw * bw)
                p.max_freq, p.min_freq = 2.0 * np.pi * 4500.0 / fs, -2.0 * np.pi * 4500.0 / fs
                p.phase, p.freq, p.locksig = phase, freq, 1.0
                first = start + pre
            oracle32.lib.pdto_chain_feed(c, iq[2 * first: 2 * stop].ctypes.data, stop - first)
            ln = C.c_size_t(0)
            text = C.string_at(oracle32.lib.pdto_chain_text(c, C.byref(ln)), ln.value).decode()
        finally:
            oracle32.lib.pdto_chain_free(c)
        stitched += [(g, b) for g, b in _to_samples(_parse(text, 0.0), axis, first) if lo <= g < hi]
    assert [b for _, b in stitched] == [b for _, b in serial]
    # positions: the same sync bit within one symbol (the printed time has 1e-5 s resolution; each segment restarts the
    # reference's float time axis, which is why ownership is decided in samples, not in the drifting time column)
    sps = max(round(150000.0 / fs), 1) * fs / 16640.3
    tol = sps + 3e-5 * fs
    assert all(abs(a - b) <= tol for (a, _), (b, _) in zip(stitched, serial))
    gs = [g for g, _ in stitched]
    assert all(0.09 * fs < b - a < 0.11 * fs for a, b in zip(gs, gs[1:]))                      # one frame per 0.1 s, none twice
