#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

Inputs  : /root/reference (read-only) + oracle/_ref/* (reference compiled by oracle/Makefile).
Outputs : small fixtures committed under tests/golden/ — the GPU box has no /root/reference, so
          every `-m gpu` test, smoke() and bench.py read only these.

  5sec_clip.wav, argos_401650kHz.wav      input captures (data fixtures, byte copies)
  poes_5sec_clip_frames.txt               demodPOES_ref output (minorFrames_*.txt) on 5sec_clip.wav
  argos_packets.txt                       demodARGOS_ref output (packets_*.txt)
  poes_minorFrame_bundled.txt             the reference's own (older-build) golden POESTIPdemod/minorFrame.txt
  argos_packets_bundled.txt               the reference's own ARGOSdemod/packets.txt
  bytesync_kat.json                       the known-answer bit strings of POESTIPdemod/ByteSync.c:8,10 with
                                          the frame counts / sync positions the reference library yields
  stage_vectors_f32.npz / _f64.npz        seeded per-stage input/output vectors produced by calling the
                                          reference .so through ctypes (one fresh dlopen per stage)
  synth_poes_c2_small.npz                 a seeded synthetic POES capture @250 ksps + the reference's frames
"""
import json
import os
import re
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import pyoracle as po  # noqa: E402

REF = os.environ.get("PDT_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
TWO_PI = 2.0 * np.pi


def main():
    po.build(ref=True)
    shutil.copy(f"{REF}/5sec_clip.wav", f"{OUT}/5sec_clip.wav")
    shutil.copy(f"{REF}/ARGOSdemod/12-21-56_401650kHz.wav", f"{OUT}/argos_401650kHz.wav")
    shutil.copy(f"{REF}/POESTIPdemod/minorFrame.txt", f"{OUT}/poes_minorFrame_bundled.txt")
    shutil.copy(f"{REF}/ARGOSdemod/packets.txt", f"{OUT}/argos_packets_bundled.txt")
    os.chmod(f"{OUT}/5sec_clip.wav", 0o644)
    os.chmod(f"{OUT}/argos_401650kHz.wav", 0o644)

    so, txt = po.run_ref_cli("POES", f"{REF}/5sec_clip.wav")
    open(f"{OUT}/poes_5sec_clip_frames.txt", "w").write(txt)
    m = re.search(r"PLL locked at (-?[\d.]+)Hz", so)
    tail = so.replace("\r", "\n").strip().splitlines()
    stats = re.search(r"(\d+) Sym : (\d+) Bits : (\d+) Frames", so.replace("\r", "\n").split("That took")[0].splitlines()[-1])
    meta = {"poes_lock_hz": float(m.group(1)), "poes_symbols": int(stats.group(1)), "poes_bits": int(stats.group(2)),
            "poes_frames": int(stats.group(3)),
            "poes_norm_factor": float(re.search(r"Normalization Factor: ([\d.]+)", so).group(1))}
    so, txt = po.run_ref_cli("ARGOS", f"{REF}/ARGOSdemod/12-21-56_401650kHz.wav")
    open(f"{OUT}/argos_packets.txt", "w").write(txt)
    meta["argos_lock_hz"] = float(re.search(r"PLL locked at (-?[\d.]+)Hz", so).group(1))
    meta["argos_norm_factor"] = float(re.search(r"Normalization Factor: ([\d.]+)", so).group(1))
    st = re.findall(r"(\d+) Sym : (\d+) Bits : (\d+) Packets", so)[-1]
    meta.update(argos_symbols=int(st[0]), argos_bits=int(st[1]), argos_packets=int(st[2]))
    json.dump(meta, open(f"{OUT}/cli_meta.json", "w"), indent=1)

    # ---- ByteSync KAT strings (POESTIPdemod/ByteSync.c:8 and :10) --------------------------------
    src = open(f"{REF}/POESTIPdemod/ByteSync.c").read().splitlines()
    kats = {}
    for name, line in (("kat_line8", src[7]), ("kat_line10", src[9])):
        bits = re.search(r'"([01]+)"', line).group(1)
        ref = po.RefLib("f32")
        arr = np.frombuffer(bits.encode(), np.uint8)
        n = ref.bytesync(arr)
        text = ref.bytesync_text()
        kats[name] = {"bits": bits, "frames": int(n), "text": text}
    json.dump(kats, open(f"{OUT}/bytesync_kat.json", "w"))

    # ---- per-stage vectors through the reference .so ---------------------------------------------
    for prec in ("f32", "f64"):
        np.savez_compressed(f"{OUT}/stage_vectors_{prec}.npz", **stage_vectors(prec))

    # ---- small synthetic C2-shaped capture with the reference's answer ---------------------------
    try:
        from tests.synth_ref import make_poes_capture  # numpy generator shared with the tests
    except Exception as e:  # pragma: no cover
        print("synthetic golden skipped:", e)
        return
    pcm, info = make_poes_capture(n_samples=400_000, fs=250_000, seed=7, esn0_db=12.0, doppler_hz=-1234.0)
    wav = f"{OUT}/_tmp_synth.wav"
    write_wav(wav, 250_000, pcm)
    so, txt = po.run_ref_cli("POES", wav)
    os.remove(wav)
    np.savez_compressed(f"{OUT}/synth_poes_c2_small.npz", pcm=pcm, frames_text=np.array(txt), fs=250_000,
                        seed=7, n_frames_sent=info["n_frames"])
    print("synthetic golden:", len(txt.splitlines()), "frames decoded of", info["n_frames"], "sent")


def write_wav(path, rate, pcm_iq_int16):
    import struct
    data = np.ascontiguousarray(pcm_iq_int16, np.int16).tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 2, rate, rate * 4, 4, 16)
    hdr += b"data" + struct.pack("<I", len(data))
    open(path, "wb").write(hdr + data)


def stage_vectors(prec):
    """Seeded inputs pushed through each reference function in two chunks (state carry-over)."""
    rng = np.random.default_rng(1234)
    dt = np.float32 if prec == "f32" else np.float64
    v = {}
    Fs = 50000.0 if prec == "f32" else 5000.0
    n = 6000
    t = np.arange(2 * n)
    # a PM carrier at -900 Hz (f32) / -120 Hz (f64) with noise: realistic enough to lock
    fo = -900.0 if prec == "f32" else -120.0
    sym = np.repeat(rng.integers(0, 2, 2 * n // 6 + 1) * 2 - 1, 6)[: 2 * n]
    ph = TWO_PI * fo * t / Fs + 0.4 + 1.169 * sym
    iq = (0.05 * np.exp(1j * ph) + 0.01 * (rng.standard_normal(2 * n) + 1j * rng.standard_normal(2 * n)))
    iq = np.stack([iq.real, iq.imag], -1).astype(dt).reshape(-1)
    v["iq"] = iq
    ref = po.RefLib(prec)
    v["static_gain"] = np.array(ref.static_gain(iq[: 2 * n]))
    if prec == "f32":
        args = (Fs, 4500.0, 0.08, 0.3979 * (TWO_PI / np.float32(Fs)), 127.3240 * (TWO_PI / np.float32(Fs)),
                10.3451 * (TWO_PI / np.float32(Fs)))
    else:
        args = (Fs, 550.0, 0.1, 3.1831 * (TWO_PI / Fs), 16 * (TWO_PI / Fs), 16 * (TWO_PI / Fs))
    o1, l1, a1 = ref.pll(iq[: 2 * n], *args, want_lock=True)
    o2, l2, a2 = ref.pll(iq[2 * n:], *args, want_lock=True)
    v["pll_args"] = np.array(args, np.float64)
    v["pll_out"] = np.concatenate([o1, o2]); v["pll_lock"] = np.concatenate([l1, l2]); v["pll_avg"] = np.array([a1, a2])
    # FIR
    x = rng.standard_normal(2 * n).astype(dt)
    v["fir_x"] = x
    if prec == "f32":
        for L in (1, 3, 8):
            ref = po.RefLib(prec)
            h = ref.make_lpfir(26 * L, 11000.0, np.float32(150000.0), L)
            tin = np.arange(2 * n + 1, dtype=dt)
            ya, _ = ref.fir_interp(tin[: n + 1], x[:n], h, L)
            yb, _ = ref.fir_interp(tin[n:], x[n:], h, L)
            v[f"h_L{L}"] = h; v[f"fir_interp_L{L}"] = np.concatenate([ya, yb])
    ref = po.RefLib(prec)
    h = ref.make_lpfir(50, 700.0, 5000.0, 1)
    v["h_argos"] = h
    v["fir_plain"] = np.concatenate([ref.fir(x[:n], h), ref.fir(x[n:], h)])
    # AGC (incl. strong-signal region to exercise the attack branch and the clamps)
    xa = (x * np.where(np.arange(2 * n) % 4000 < 2000, 0.05, 3.0)).astype(dt)
    v["agc_x"] = xa
    v["agc_y"] = np.concatenate([ref.agc(xa[:n], 17.5, 0.0033, 0.0067), ref.agc(xa[n:], 17.5, 0.0033, 0.0067)])
    # Gardner on a smooth split-phase-like waveform, chunked, with the stale-read padding zeroed
    sps = 9.014 if prec == "f32" else 6.25
    w = np.sin(np.pi * np.arange(2 * n) / sps + 0.3) + 0.1 * rng.standard_normal(2 * n)
    v["gar_x"] = w.astype(dt)
    buf = np.zeros(n + 16, dt)
    syms, idxs = [], []
    FsI, baud = (150000, 16640.3) if prec == "f32" else (5000, 800.0)
    for k in range(2):
        buf[:n] = w[k * n:(k + 1) * n]
        s, i = ref.gardner(buf, n, FsI, baud, 0.1, 3.0)
        syms.append(s); idxs.append(i + k * n)
    v["gar_sym"] = np.concatenate(syms); v["gar_idx"] = np.concatenate(idxs).astype(np.int64)
    v["gar_args"] = np.array([FsI, baud, 0.1, 3.0])
    # Manchester
    ms = (rng.standard_normal(4001) * 1.5).astype(dt)
    v["man_sym"] = ms
    thr = 1.0 if prec == "f32" else 0.5
    v["man_bits"] = np.concatenate([ref.manchester(ms[:1777], thr), ref.manchester(ms[1777:], thr)])
    return v


if __name__ == "__main__":
    main()
