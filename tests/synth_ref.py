"""Seeded synthetic POES-TIP / ARGOS IQ capture generators (numpy; test + bench infrastructure).

Signal model follows SURVEY.md §8(d):
  TIP minor frame = 104 bytes: ED E2 | spacecraft id (top three bits 0 => the 19-bit sync 1110110111100010000) |
  … bytes 4-5 carry the 9-bit minor-frame counter (wraps at 320, standalone_matlab/Functionized/daytimeDecode.m:4)
  … byte 103 carries five even-parity bits (checkParity.m:20-90).
  bits -> split-phase chips: '1' = (+,-), '0' = (-,+)  (ManchesterDecode.c:60-82), 16640.3 chips/s (main.c:90)
  x[n] = A·exp(j(2π(f0·t + ½·drift·t²) + θ0 + m·chip[n])) + σ·(N(0,1)+jN(0,1))/√2,  m = 1.169 rad
  quantised to int16 like a WAV capture (wave.c normalises by 32768).
"""
from __future__ import annotations

import numpy as np

MOD_INDEX = 1.169
POES_CHIP_RATE = 8320 * 2 + 0.3
FRAME_BYTES = 104


def _popcount_bytes(b: np.ndarray) -> int:
    return int(np.unpackbits(np.asarray(b, np.uint8)).sum())


def tip_frames(n_frames: int, seed: int, spacecraft: int = 8, counter0: int = 0) -> np.ndarray:
    """[n_frames,104] uint8 minor frames with valid counter and parity word."""
    rng = np.random.default_rng(seed)
    fr = rng.integers(0, 256, (n_frames, FRAME_BYTES), dtype=np.uint8)
    fr[:, 0], fr[:, 1], fr[:, 2] = 0xED, 0xE2, spacecraft
    cnt = (counter0 + np.arange(n_frames)) % 320
    fr[:, 4] = (fr[:, 4] & 0xFE) | (cnt >> 8).astype(np.uint8)
    fr[:, 5] = (cnt & 0xFF).astype(np.uint8)
    # an accidental sync word inside the payload is ignored by ByteSync while in-frame; nothing to do.
    for f in fr:
        w = int(f[103]) & 0xC1                     # keep CPU A/B flags and bit 0
        for k, (lo, hi) in enumerate(((2, 18), (19, 35), (36, 52), (53, 69), (70, 86))):
            w |= (_popcount_bytes(f[lo:hi + 1]) & 1) << (5 - k)
        f[103] = w
    return fr


def frames_to_bits(frames: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.asarray(frames, np.uint8).reshape(-1))


def check_parity(frame: np.ndarray) -> bool:
    """checkParity.m:20-90 on one 104-byte frame."""
    w = int(frame[103])
    for k, (lo, hi) in enumerate(((2, 18), (19, 35), (36, 52), (53, 69), (70, 86))):
        if (_popcount_bytes(frame[lo:hi + 1]) & 1) != ((w >> (5 - k)) & 1):
            return False
    return True


def frame_counter(frame: np.ndarray) -> int:
    return ((int(frame[4]) & 1) << 8) | int(frame[5])


def parse_frames_text(text: str):
    """minorFrames_*.txt / packets_*.txt -> list of (time_str, inverse, bytes[np.uint8]) ; partial rows kept."""
    out = []
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        t = tok[0]
        inv = t.endswith("i")
        out.append((t.rstrip("i"), inv, np.array([int(x, 16) for x in tok[1:]], np.uint8)))
    return out


def make_poes_capture(n_samples: int, fs: float, seed: int, esn0_db: float = 12.0, doppler_hz: float = 1000.0,
                      drift_hz_s: float = 20.0, amplitude: float = 0.25, theta0: float = 0.7,
                      lead_in_s: float = 0.0, spacecraft: int = 8, counter0: int | None = None,
                      chip_rate: float = POES_CHIP_RATE, frames_hook=None):
    """Returns (pcm int16 [2n] interleaved I,Q ; info dict).  `frames_hook(frames)` may edit the [n,104] frame table in place
    before it is modulated (crafted payloads)."""
    rng = np.random.default_rng(seed)
    sps = fs / chip_rate
    n_chips = int(np.ceil(n_samples / sps)) + 4
    n_frames = n_chips // (2 * 8 * FRAME_BYTES) + 4
    if counter0 is None:
        counter0 = int(rng.integers(0, 320))
    frames = tip_frames(n_frames, seed + 1000003, spacecraft, counter0)
    if frames_hook is not None:
        frames_hook(frames)
    bits = frames_to_bits(frames)
    # random frame phase so that captures do not all start on a frame boundary (counter stays continuous)
    start = int(rng.integers(0, 2 * 8 * FRAME_BYTES))
    bits = bits[start:]
    chips = np.empty(2 * bits.size, np.int8)
    chips[0::2] = np.where(bits == 1, 1, -1)
    chips[1::2] = -chips[0::2]
    t = np.arange(n_samples, dtype=np.float64) / fs
    chip_idx = np.floor(np.arange(n_samples, dtype=np.float64) / sps).astype(np.int64)
    d = chips[chip_idx].astype(np.float64)
    if lead_in_s > 0:
        d[: int(lead_in_s * fs)] = 0.0
    phase = 2 * np.pi * (doppler_hz * t + 0.5 * drift_hz_s * t * t) + theta0 + MOD_INDEX * d
    esn0 = 10.0 ** (esn0_db / 10.0)
    sigma = amplitude * np.sin(MOD_INDEX) * np.sqrt(sps / esn0)
    x = amplitude * np.exp(1j * phase)
    x += sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples)) / np.sqrt(2.0)
    pcm = np.empty(2 * n_samples, np.int16)
    pcm[0::2] = np.clip(np.rint(x.real * 32768.0), -32768, 32767).astype(np.int16)
    pcm[1::2] = np.clip(np.rint(x.imag * 32768.0), -32768, 32767).astype(np.int16)
    return pcm, dict(n_frames=int(n_samples / sps / (2 * 8 * FRAME_BYTES)), frames=frames, bit_start=start,
                     sigma=sigma, sps=sps, counter0=counter0)


def make_argos_capture(n_samples: int, fs: float = 5000.0, seed: int = 0, n_bursts: int = 4, snr_db: float = 15.0,
                       doppler_hz: float = -200.0, amplitude: float = 0.2):
    """ARGOS-like bursts: 160 ms carrier + 15 ones + 00010111 + 1 + 0000 + payload at 400 bps
    split-phase ±1.1 rad, noise between bursts.  Returns (pcm int16 [2n], info)."""
    rng = np.random.default_rng(seed)
    chip_rate = 800.0
    sps = fs / chip_rate
    x = np.zeros(n_samples, np.complex128)
    payloads = []
    spacing = n_samples // (n_bursts + 1)
    for b in range(n_bursts):
        pay = rng.integers(0, 2, 7 * 8 + 24).astype(np.uint8)
        bits = np.concatenate([np.ones(15, np.uint8), np.array([0, 0, 0, 1, 0, 1, 1, 1, 1, 0, 0, 0, 0], np.uint8), pay])
        chips = np.empty(2 * bits.size)
        chips[0::2] = np.where(bits == 1, 1.0, -1.0)
        chips[1::2] = -chips[0::2]
        n_car = int(0.160 * fs)
        n_dat = int(chips.size * sps)
        d = np.concatenate([np.zeros(n_car), chips[np.floor(np.arange(n_dat) / sps).astype(int)]])
        s0 = spacing * (b + 1)
        m = min(d.size, n_samples - s0)
        tt = np.arange(m) / fs
        fo = doppler_hz + 30.0 * b
        x[s0:s0 + m] = amplitude * np.exp(1j * (2 * np.pi * fo * tt + 0.3 * b + 1.1 * d[:m]))
        payloads.append(pay)
    sigma = amplitude / np.sqrt(10.0 ** (snr_db / 10.0))
    x += sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples)) / np.sqrt(2.0)
    pcm = np.empty(2 * n_samples, np.int16)
    pcm[0::2] = np.clip(np.rint(x.real * 32768.0), -32768, 32767).astype(np.int16)
    pcm[1::2] = np.clip(np.rint(x.imag * 32768.0), -32768, 32767).astype(np.int16)
    return pcm, dict(payloads=payloads)
