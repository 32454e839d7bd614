"""CPU oracle (oracle/pdt_oracle.c) against the committed golden vectors.

The goldens were produced by the UNMODIFIED reference (tests/golden/make_golden.py); these tests need
neither /root/reference nor oracle/_ref, so they also pin the oracle on the GPU box.
"""
import json
import os

import numpy as np
import pytest

import pyoracle as po
from tests.synth_ref import check_parity, frame_counter, parse_frames_text

TWO_PI = 2.0 * np.pi


def _golden(golden_dir, name):
    return os.path.join(golden_dir, name)


def test_poes_5sec_clip_frames_bit_exact(oracle32, golden_dir):
    rate, pcm = po.read_wav_pcm16(_golden(golden_dir, "5sec_clip.wav"))
    res = oracle32.chain(oracle32.pcm16_to_complex(pcm), rate)
    want = open(_golden(golden_dir, "poes_5sec_clip_frames.txt")).read()
    assert res["text"] == want                     # bytes AND the float-accumulated time column
    meta = json.load(open(_golden(golden_dir, "cli_meta.json")))
    assert res["total_symbols"] == meta["poes_symbols"]
    assert res["total_bits"] == meta["poes_bits"]
    assert res["total_frames"] == meta["poes_frames"]
    assert f"{res['lock_freq_hz']:.2f}" == f"{meta['poes_lock_hz']:.2f}"
    assert f"{res['norm_factor']:.6f}" == f"{meta['poes_norm_factor']:.6f}"


def test_poes_frames_selfcheck(golden_dir):
    """Counter continuity 275…319,0,1,2 and spacecraft id 0x08 (SURVEY §8c self-checks)."""
    fr = parse_frames_text(open(_golden(golden_dir, "poes_5sec_clip_frames.txt")).read())
    full = [f for f in fr if f[2].size == 104]
    assert len(fr) == 48 and len(full) == 47
    cnt = [frame_counter(f[2]) for f in full]
    assert cnt[0] == 275 and all((b - a) % 320 == 1 for a, b in zip(cnt, cnt[1:]))
    assert all(f[2][2] == 0x08 for f in full)
    assert sum(check_parity(f[2]) for f in full) >= 40


def test_poes_vs_bundled_older_golden(golden_dir):
    """POESTIPdemod/minorFrame.txt (older build): rows line up with HEAD frames 2…48 except 2 bytes."""
    head = parse_frames_text(open(_golden(golden_dir, "poes_5sec_clip_frames.txt")).read())
    old = parse_frames_text(open(_golden(golden_dir, "poes_minorFrame_bundled.txt")).read())
    assert len(old) == len(head) - 1
    diffs = 0
    for o, h in zip(old, head[1:]):
        n = min(o[2].size, h[2].size)
        diffs += int((o[2][:n] != h[2][:n]).sum())
    assert diffs == 2


def test_argos_packets_bit_exact(oracle64, golden_dir):
    rate, pcm = po.read_wav_pcm16(_golden(golden_dir, "argos_401650kHz.wav"))
    res = oracle64.chain(oracle64.pcm16_to_complex(pcm), rate, argos=True)
    assert res["text"] == open(_golden(golden_dir, "argos_packets.txt")).read()
    meta = json.load(open(_golden(golden_dir, "cli_meta.json")))
    assert (res["total_symbols"], res["total_bits"], res["total_frames"]) == \
        (meta["argos_symbols"], meta["argos_bits"], meta["argos_packets"])
    # the reference's bundled packets.txt (older format) pins the first six payload bytes
    old = parse_frames_text(open(_golden(golden_dir, "argos_packets_bundled.txt")).read())
    new = parse_frames_text(res["text"])
    for o, n in zip(old, new):
        assert list(o[2][2:8]) == list(n[2][:6])
        assert abs(float(o[0]) - float(n[0])) < 1e-3


@pytest.mark.parametrize("name,frames", [("kat_line8", 23), ("kat_line10", 4)])
def test_bytesync_kat(oracle32, golden_dir, name, frames):
    kat = json.load(open(_golden(golden_dir, "bytesync_kat.json")))[name]
    assert kat["frames"] == frames
    st = oracle32.new_state("bytesync")
    bits = np.frombuffer(kat["bits"].encode(), np.uint8)
    # feed in ragged pieces: state must carry across calls
    n = 0
    for lo, hi in ((0, 1), (1, 777), (777, 778), (778, bits.size)):
        n += oracle32.bytesync(st, bits[lo:hi], "poes", time=np.zeros(hi - lo + 1, np.float32))
    assert n == frames
    assert oracle32.bytesync_text(st) == kat["text"]


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_stage_vectors(prec, golden_dir):
    o = po.Oracle(prec)
    v = np.load(_golden(golden_dir, f"stage_vectors_{prec}.npz"))
    dt = o.dt
    iq = v["iq"]
    n = iq.size // 4
    assert o.static_gain(iq[: 2 * n]) == float(v["static_gain"])
    a = [float(x) for x in v["pll_args"]]
    st = o.new_state("pll")
    o1, l1, a1, _, _ = o.pll(st, iq[: 2 * n], *a, want_lock=True)
    o2, l2, a2, _, _ = o.pll(st, iq[2 * n:], *a, want_lock=True)
    assert np.array_equal(np.concatenate([o1, o2]), v["pll_out"])
    assert np.array_equal(np.concatenate([l1, l2]), v["pll_lock"])
    assert [a1, a2] == list(v["pll_avg"])
    x = v["fir_x"]
    if prec == "f32":
        for L in (1, 3, 8):
            h = o.make_lpfir(26 * L, 11000.0, np.float32(150000.0), L)
            assert np.array_equal(h, v[f"h_L{L}"])
            st = o.new_state("fir")
            tin = np.arange(2 * n + 1, dtype=dt)
            ya, _ = o.fir_interp(st, tin[: n + 1], x[:n], h, L)
            yb, _ = o.fir_interp(st, tin[n:], x[n:], h, L)
            assert np.array_equal(np.concatenate([ya, yb]), v[f"fir_interp_L{L}"])
    h = o.make_lpfir(50, 700.0, 5000.0, 1)
    assert np.array_equal(h, v["h_argos"])
    st = o.new_state("fir")
    assert np.array_equal(np.concatenate([o.fir(st, x[:n], h), o.fir(st, x[n:], h)]), v["fir_plain"])
    st = o.new_state("agc")
    ya, _ = o.agc(st, v["agc_x"][:n], 17.5, 0.0033, 0.0067)
    yb, _ = o.agc(st, v["agc_x"][n:], 17.5, 0.0033, 0.0067)
    assert np.array_equal(np.concatenate([ya, yb]), v["agc_y"])
    st = o.new_state("gardner")
    FsI, baud, rng_, kp = v["gar_args"]
    buf = np.zeros(n + 16, dt)
    syms, idxs = [], []
    for k in range(2):
        buf[:n] = v["gar_x"][k * n:(k + 1) * n]
        s, i, _ = o.gardner(st, buf, n, int(FsI), float(baud), float(rng_), float(kp))
        syms.append(s)
        idxs.append(i.astype(np.int64) + k * n)
    assert np.array_equal(np.concatenate(syms), v["gar_sym"])
    assert np.array_equal(np.concatenate(idxs), v["gar_idx"])
    st = o.new_state("manchester")
    thr = 1.0 if prec == "f32" else 0.5
    bits = np.concatenate([o.manchester(st, v["man_sym"][:1777], thr), o.manchester(st, v["man_sym"][1777:], thr)])
    assert np.array_equal(bits, v["man_bits"])


def test_synth_c2_small(oracle32, golden_dir):
    g = np.load(_golden(golden_dir, "synth_poes_c2_small.npz"))
    res = oracle32.chain(oracle32.pcm16_to_complex(g["pcm"]), int(g["fs"]))
    assert res["text"] == str(g["frames_text"])
    assert res["L"] == 1 and res["N"] == 26
    full = [f for f in parse_frames_text(res["text"]) if f[2].size == 104]
    assert len(full) >= int(g["n_frames_sent"]) - 1
    assert all(check_parity(f[2]) for f in full)


def test_l0_emits_nothing(oracle32):
    """Fs >= 300 ksps -> L = rint(150000/Fs) = 0 -> the reference silently produces no output."""
    rng = np.random.default_rng(0)
    iq = (0.1 * rng.standard_normal(40000)).astype(np.float32)
    res = oracle32.chain(iq, 2_000_000)
    assert res["L"] == 0 and res["total_symbols"] == 0 and res["text"] == ""
    res = oracle32.chain(iq, 2_000_000, force_min_L1=True)
    assert res["L"] == 1 and res["total_symbols"] > 0
