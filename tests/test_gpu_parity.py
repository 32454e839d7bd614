"""GPU parity tests proper (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI of
libpdt_f32.so / libpdt_f64.so and is compared with the CPU oracle on the same inputs, and with the golden
vectors generated from the unmodified reference.  /root/reference is NOT needed here.

Tolerances: integer / byte / index outputs are bit-exact.  The float (POES) exact-order path is bit-exact at
every stage (the kernels restate glibc's sinf/cosf and never contract FMAs).  The double (ARGOS) path uses
CUDA's sin/cos (<= 2 ulp from glibc), so its PLL output is compared with rtol 1e-12 and everything decided
downstream (symbol picks, bits, packets) is bit-exact on the fixtures.
"""
import ctypes as C
import importlib
import json
import os
import subprocess

import numpy as np
import pytest

import pyoracle as po
from tests.synth_ref import check_parity, frame_counter, make_argos_capture, make_poes_capture, parse_frames_text

pdt = importlib.import_module("project-desert-tortoise_b200")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TWO_PI = 2.0 * np.pi


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


def _golden(golden_dir, name):
    return os.path.join(golden_dir, name)


def _frames_text_equal_bytes(text_a, text_b):
    a, b = parse_frames_text(text_a), parse_frames_text(text_b)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[1] == y[1] and np.array_equal(x[2], y[2])


class DevTraces:
    """Device trace taps for one capture, allocated with torch."""

    def __init__(self, torch, dt, n, L, cap):
        self.torch, self.dt = torch, dt
        tdt = torch.float32 if dt == np.float32 else torch.float64
        z = lambda m, t=tdt: torch.zeros(m, dtype=t, device="cuda")
        self.t = dict(pll_phase=z(n), pll_freq=z(n), pll_out=z(n), lock=z(n), lpf=z(n * L), agc=z(n * L), sym=z(cap),
                      gardner_err=z(cap), gardner_idx=z(cap, torch.int64), bits=z(cap, torch.uint8))
        self.cap = cap

    def struct(self, skip=()):
        s = pdt.Traces()
        for k, v in self.t.items():
            if k in skip:
                continue
            setattr(s, k, v.data_ptr())
        s.cap = self.cap
        return s

    def host(self, k, n=None):
        a = self.t[k].cpu().numpy()
        return a if n is None else a[:n]


ENGINES = {"exact": pdt.PDT_ENGINE_EXACT, "tiled": pdt.PDT_ENGINE_TILED}


def _run_batch_with_traces(torch, prec, mode, fs, iq, chunk=None, force_l1=False, engine="exact", **overrides):
    p = pdt.default_params(prec, mode, fs)
    if chunk:
        p.chunk = chunk
    p.force_min_interp1 = int(force_l1)
    p.engine = ENGINES[engine]
    for k, v in overrides.items():
        setattr(p, k, v)
    dt = np.float32 if prec == "f32" else np.float64
    iq = np.ascontiguousarray(iq, dt)
    n = iq.size // 2
    d = pdt.Demod(prec, p, 1, n, 128)
    L = max(d.params.interp, 1)
    cap = n * L // 4 + 64
    tr = DevTraces(torch, dt, n, L, cap)
    d_iq = torch.from_numpy(iq).cuda()
    assert d.engine == ENGINES[engine]
    # the per-sample frequency and lock-detector taps exist only in the exact engine (asking for them selects it)
    d.demod_device(d_iq.data_ptr(), 1, n, traces=[tr.struct(skip=("pll_freq", "lock") if engine == "tiled" else ())])
    stats, frames = d.fetch(1)
    return d, stats[0], frames[0], tr


# ------------------------------------------------------------------------------------------------------
# whole chain, batch API, with every intermediate stream compared against the oracle
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ["exact", "tiled"])
def test_poes_5sec_clip_bit_exact_all_stages(torch_cuda, oracle32, golden_dir, engine):
    rate, pcm = po.read_wav_pcm16(_golden(golden_dir, "5sec_clip.wav"))
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, rate, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, rate, iq, engine=engine)
    assert (st["n_symbols"], st["n_bits"], st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert st["locked"] == 1 and st["lock_sample"] == want["lock_sample"]
    assert np.float32(st["norm_factor"]) == np.float32(want["norm_factor"])
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    for k_dev, k_or in (("pll_phase", "tr_phase"), ("pll_freq", "tr_freq"), ("pll_out", "tr_pll_out"), ("lpf", "tr_lpf"),
                        ("agc", "tr_agc")):
        if engine == "tiled" and k_dev == "pll_freq":
            continue
        assert np.array_equal(tr.host(k_dev), want[k_or]), k_dev
    assert np.array_equal(tr.host("sym", ns), want["tr_sym"])
    assert np.array_equal(tr.host("gardner_err", ns), want["tr_gerr"])
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    # frame bytes + the time column against the file the unmodified reference wrote
    text = d.format_frames(fr, int(st["n_frames"]))
    golden = open(_golden(golden_dir, "poes_5sec_clip_frames.txt")).read()
    _frames_text_equal_bytes(text, golden)
    assert text == golden
    meta = json.load(open(_golden(golden_dir, "cli_meta.json")))
    assert f"{st['lock_freq_hz']:.2f}" == f"{meta['poes_lock_hz']:.2f}"


def test_poes_pcm16_ingest_equals_float_ingest(torch_cuda, golden_dir):
    """int16 PCM straight from the WAV data chunk (4 B/sample) must give the same frames as normalised floats."""
    rate, pcm = po.read_wav_pcm16(_golden(golden_dir, "5sec_clip.wav"))
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, rate)
    d = pdt.Demod("f32", p, 1, pcm.size // 2, 64)
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    text = d.format_frames(fr[0], int(st[0]["n_frames"]))
    assert text == open(_golden(golden_dir, "poes_5sec_clip_frames.txt")).read()


def test_argos_wav_packets(torch_cuda, oracle64, golden_dir):
    rate, pcm = po.read_wav_pcm16(_golden(golden_dir, "argos_401650kHz.wav"))
    iq = oracle64.pcm16_to_complex(pcm)
    want = oracle64.chain(iq, rate, argos=True, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f64", pdt.PDT_MODE_ARGOS, rate, iq)
    text = d.format_frames(fr, int(st["n_frames"]))
    golden = open(_golden(golden_dir, "argos_packets.txt")).read()
    _frames_text_equal_bytes(text, golden)
    assert text == golden
    assert (st["n_symbols"], st["n_bits"], st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert st["lock_sample"] == want["lock_sample"]
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    # double path: CUDA sin/cos vs glibc differ by <= 2 ulp -> tolerance on the analogue streams, exact decisions
    np.testing.assert_allclose(tr.host("pll_out"), want["tr_pll_out"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(tr.host("pll_phase"), want["tr_phase"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(tr.host("agc"), want["tr_agc"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])


@pytest.mark.parametrize("fs,chunk,seed,engine", [(250000, 10000, 7, "exact"), (250000, 4096, 8, "exact"), (50000, 10000, 9, "exact"),
                                                  (18750, 10000, 10, "exact"), (250000, 10000, 7, "tiled"), (250000, 4096, 8, "tiled"),
                                                  (50000, 10000, 9, "tiled"), (75000, 10000, 11, "tiled"), (37500, 7000, 12, "tiled"),
                                                  (30000, 10000, 13, "tiled"), (25000, 10000, 15, "tiled"), (21430, 5000, 14, "tiled"),
                                                  (18750, 10000, 10, "tiled")])
def test_poes_synthetic_vs_oracle(torch_cuda, oracle32, fs, chunk, seed, engine):
    """L = 1 ... 8 (8 = the historical 8x interpolator), ragged last chunk, non-default chunk length, both engines."""
    pcm, info = make_poes_capture(int(1.3 * fs) + 123, fs, seed, esn0_db=11.0, doppler_hz=-2000.0 + 300 * seed)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, chunk=chunk, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, chunk=chunk, engine=engine)
    assert d.params.interp == {250000: 1, 75000: 2, 50000: 3, 37500: 4, 30000: 5, 25000: 6, 21430: 7, 18750: 8}[fs]
    assert want["total_frames"] >= 8
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    assert (ns, nb, st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert np.array_equal(tr.host("pll_out"), want["tr_pll_out"])
    assert np.array_equal(tr.host("agc"), want["tr_agc"])
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    text = d.format_frames(fr, int(st["n_frames"]))
    _frames_text_equal_bytes(text, want["text"])
    full = [f for f in parse_frames_text(text) if f[2].size == 104]
    cnt = [frame_counter(f[2]) for f in full]
    if fs >= 50000:        # at 18.75 ksps (1.13 samples per chip before the 8x interpolator) the link itself makes bit errors
        assert all((b - a) % 320 == 1 for a, b in zip(cnt, cnt[1:]))
        assert sum(check_parity(f[2]) for f in full) >= len(full) - 1


def test_tiled_long_capture_many_tiles_bit_exact(torch_cuda, oracle32):
    """4 s @ 250 ksps: ~30 PLL tiles and several AGC tiles; every stream bit-identical to the serial oracle, and the
    speculation is almost always accepted (re-runs are allowed, wrong results are not)."""
    fs = 250000
    pcm, info = make_poes_capture(1_000_000, fs, 21, esn0_db=14.0, doppler_hz=700.0, drift_hz_s=45.0, amplitude=0.2)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, engine="tiled")
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    assert st["locked"] == 1 and st["lock_sample"] == want["lock_sample"]
    for k_dev, k_or in (("pll_phase", "tr_phase"), ("pll_out", "tr_pll_out"), ("lpf", "tr_lpf"), ("agc", "tr_agc")):
        bad = np.nonzero(tr.host(k_dev) != want[k_or])[0]
        assert bad.size == 0, (k_dev, bad[:5], bad.size)
    assert (ns, nb, st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("gardner_err", ns), want["tr_gerr"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    _frames_text_equal_bytes(d.format_frames(fr, int(st["n_frames"])), want["text"])
    pll_rerun, agc_rerun, acq_restarts, tiles = d.tiled_counters()
    print("tiled counters", pll_rerun, agc_rerun, acq_restarts, tiles)
    assert tiles >= 12 and pll_rerun <= 3 and agc_rerun <= 3


@pytest.mark.parametrize("fs,n,seed", [(50000, 600_000, 31), (18750, 230_000, 32)])
def test_tiled_long_interpolating_capture_bit_exact(torch_cuda, oracle32, fs, n, seed):
    """L = 3 (12 s) and L = 8 (12 s): many PLL tiles and AGC tiles on the interpolated rate, the branch-looped FIR of k_front<L>
    over hundreds of CTAs — every stream bit-identical to the serial oracle."""
    pcm, info = make_poes_capture(n, fs, seed, esn0_db=14.0, doppler_hz=700.0, drift_hz_s=20.0, amplitude=0.2)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, engine="tiled")
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    assert d.params.interp == {50000: 3, 18750: 8}[fs] and want["total_frames"] >= 100
    assert st["locked"] == 1 and st["lock_sample"] == want["lock_sample"]
    for k_dev, k_or in (("pll_phase", "tr_phase"), ("pll_out", "tr_pll_out"), ("lpf", "tr_lpf"), ("agc", "tr_agc")):
        bad = np.nonzero(tr.host(k_dev) != want[k_or])[0]
        assert bad.size == 0, (k_dev, bad[:5], bad.size)
    assert (ns, nb, st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    _frames_text_equal_bytes(d.format_frames(fr, int(st["n_frames"])), want["text"])
    assert d.tiled_counters()[3] >= 4


def test_tiled_failed_speculation_is_repaired(torch_cuda, oracle32):
    """Warm-up windows far too short to converge: the verification must reject them and the re-run must restore the
    exact serial result (this is what makes the tiled engine exact by construction, not by luck)."""
    fs = 250000
    pcm, info = make_poes_capture(400_000, fs, 22, esn0_db=14.0, doppler_hz=-900.0, amplitude=0.3)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, engine="tiled",
                                           pll_warm=1024, pll_tile=20000, agc_min_tile=4096)
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    for k_dev, k_or in (("pll_phase", "tr_phase"), ("pll_out", "tr_pll_out"), ("agc", "tr_agc")):
        assert np.array_equal(tr.host(k_dev), want[k_or]), k_dev
    assert (ns, nb, st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    pll_rerun, agc_rerun, _, _ = d.tiled_counters()
    assert pll_rerun >= 5          # the speculation really did fail


@pytest.mark.parametrize("kind", ["early", "late", "never"])
def test_tiled_two_pass_acquisition_bit_exact(torch_cuda, oracle32, kind):
    """First acquisition pass of 40960 samples only: a capture that latches inside it, one that latches in the second
    pass (its carrier appears late) and one that never latches; every stream must equal the serial oracle."""
    fs, n = 250000, 300_000
    if kind == "never":
        rng = np.random.default_rng(11)
        pcm = (rng.standard_normal(2 * n) * 700).astype(np.int16)
    else:
        # the sweep runs to the positive frequency limit and stays there (CarrierTrackingPLL.c:232-246): a carrier that
        # appears late is only captured near +4.5 kHz
        pcm, _ = make_poes_capture(n, fs, 31, esn0_db=15.0, doppler_hz=1200.0 if kind == "early" else 4400.0, amplitude=0.25)
        if kind == "late":
            rng = np.random.default_rng(12)
            quiet = 150_000
            pcm = pcm.copy()
            pcm[: 2 * quiet] = (rng.standard_normal(2 * quiet) * 300).astype(np.int16)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, trace=True)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, engine="tiled", acq_first=40960)
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    assert st["locked"] == int(want["locked"])
    if kind == "early":
        assert st["lock_sample"] < 40960
    if kind == "late":
        assert st["locked"] == 1 and st["lock_sample"] > 40960
    if want["locked"]:
        assert st["lock_sample"] == want["lock_sample"]
    for k_dev, k_or in (("pll_phase", "tr_phase"), ("pll_out", "tr_pll_out"), ("lpf", "tr_lpf"), ("agc", "tr_agc")):
        assert np.array_equal(tr.host(k_dev), want[k_or]), k_dev
    assert (ns, nb, st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    _frames_text_equal_bytes(d.format_frames(fr, int(st["n_frames"])), want["text"])


def test_tiled_unlocked_and_tiny_captures(torch_cuda, oracle32):
    """Noise only (the PLL never latches: everything stays in the acquisition kernel), and captures shorter than one chunk."""
    fs = 250000
    rng = np.random.default_rng(5)
    noise = (rng.standard_normal(2 * 150_000) * 900).astype(np.int16)
    short, _ = make_poes_capture(5000, fs, 23)
    for pcm in (noise, short, short[:2 * 3]):
        iq = oracle32.pcm16_to_complex(pcm)
        want = oracle32.chain(iq, fs, trace=True)
        d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, engine="tiled")
        ns, nb = int(st["n_symbols"]), int(st["n_bits"])
        assert st["locked"] == int(want["locked"])
        assert (ns, nb, st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
        assert np.array_equal(tr.host("pll_out"), want["tr_pll_out"])
        assert np.array_equal(tr.host("agc"), want["tr_agc"])
        assert np.array_equal(tr.host("bits", nb), want["tr_bits"])


@pytest.mark.parametrize("acq_first", [0, 28160])
def test_tiled_stream_groups_equal_exact_engine(torch_cuda, acq_first):
    """A batch large enough to be cut into capture groups on internal streams (fork/join inside pdt_demod_device):
    stats and frames of every capture must equal the exact engine's, whatever the grouping.  With a short first
    acquisition pass the late-locking captures of every group additionally move to the slow-capture streams."""
    torch = torch_cuda
    fs, n, caps = 250000, 120_000, 200
    L = pdt.load("f32")
    d_iq = torch.empty(caps * n * 2, dtype=torch.int16, device="cuda")
    assert L.pdt_synth_poes_device(d_iq.data_ptr(), 1, caps, n, n, float(fs), 77, 0) == 0
    res = {}
    for eng in ("exact", "tiled"):
        p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
        p.engine = ENGINES[eng]
        p.acq_first = acq_first
        d = pdt.Demod("f32", p, caps, n, 16)
        d.demod_device(d_iq.data_ptr(), caps, n, pcm16=True)
        res[eng] = d.fetch(caps)
    (se, fe), (st, ft) = res["exact"], res["tiled"]
    assert se["locked"].sum() >= caps // 2 and se["n_frames"].sum() > caps
    if acq_first:      # the two-pass path must really have been exercised by both kinds of capture
        late = (se["locked"] == 0) | (se["lock_sample"] >= acq_first)
        assert late.sum() >= 3 and (~late).sum() >= 3
    for k in ("n_samples", "n_symbols", "n_bits", "n_frames", "locked", "lock_sample", "lock_freq_hz", "norm_factor",
              "final_phase", "final_freq", "final_gain", "final_next"):
        assert np.array_equal(se[k], st[k]), k
    assert np.array_equal(fe, ft)


def test_config1_single_10M_sample_stream_bit_exact_frames(torch_cuda, oracle32):
    """BASELINE configs[1]: one 10 M-sample capture @ 250 ksps (40 s, ~400 minor frames, ~300 PLL tiles, Doppler drift) on
    one GPU: counts, lock sample and every frame byte equal to the CPU oracle."""
    fs, n = 250000, 10_000_000
    pcm, info = make_poes_capture(n, fs, 77, esn0_db=13.0, doppler_hz=-1800.0, drift_hz_s=40.0, amplitude=0.15)
    want = oracle32.chain(oracle32.pcm16_to_complex(pcm), fs)
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    d = pdt.Demod("f32", p, 1, n, 512)
    assert d.engine == pdt.PDT_ENGINE_TILED
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    assert want["total_frames"] >= 390
    assert (st[0]["n_symbols"], st[0]["n_bits"], st[0]["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert st[0]["locked"] == 1 and st[0]["lock_sample"] == want["lock_sample"]
    text = d.format_frames(fr[0], int(st[0]["n_frames"]))
    _frames_text_equal_bytes(text, want["text"])
    full = [f for f in parse_frames_text(text) if f[2].size == 104]
    cnt = [frame_counter(f[2]) for f in full]
    assert all((b - a) % 320 == 1 for a, b in zip(cnt, cnt[1:]))          # no minor frame lost across ~300 tile seams
    q = d.frame_checks(1)[0]
    assert q["valid"].sum() == len(full) and q["parity_ok"][q["valid"] == 1].mean() > 0.98


@pytest.mark.parametrize("fs,total,segment,min_frames", [(250000, 10_000_000, 1_000_000, 390), (2_000_000, 12_000_000, 2_000_000, 50)])
def test_config4_stream_as_segments_matches_serial_capture(torch_cuda, fs, total, segment, min_frames):
    """BASELINE configs[4] in small (one GPU): ONE synthetic stream (Doppler crossing zero), demodulated (a) serially as a single
    capture — the path the other tests pin bit-exact against the oracle — and (b) as overlapping segments stitched by
    ownership windows (pdt_stream_*).  Acceptance: every complete minor frame of the serial result appears
    with identical bytes, sync positions within one symbol, no counter break across the seams; and a rank that holds only
    its slice of the stream produces exactly its part of the stitched list.  2 Msps uses the declared L = 1 deviation."""
    torch = torch_cuda
    stream_mod = importlib.import_module("project-desert-tortoise_b200.stream")
    L = stream_mod._bind(pdt.load("f32"))
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    if fs > 300000:
        p.force_min_interp1 = 1
    cs = torch.cuda.current_stream().cuda_stream
    d_iq = torch.empty(total * 2, dtype=torch.float32, device="cuda")
    assert L.pdt_synth_poes_stream_device(d_iq.data_ptr(), 0, 0, total, total, float(fs), 99, cs) == 0
    # (a) serial
    d = pdt.Demod("f32", p, 1, total, int(total / fs * 10) + 8)
    d.demod_device(d_iq.data_ptr(), 1, total, stream=cs)
    st, fr = d.fetch(1, cs)
    serial = fr[0][: int(st[0]["n_frames"])]
    serial_full = serial[serial["complete"] == 1]
    serial_locked = int(st[0]["locked"]) == 1
    # (b) segments
    plan = stream_mod.make_plan("f32", d.params, total, segment)
    k = plan.n_segments
    assert k >= 4
    sd = stream_mod.StreamDemod("f32", p, plan, 0, k)
    assert (sd.start, sd.n_slice) == (0, total)
    sd.run_device(d_iq.data_ptr(), stream=cs)
    s_st, s_fr = sd.fetch(cs)
    assert np.all(s_st["locked"][1:] == 1)                                   # segments behind the first start pre-locked
    out = sd.stitch_local(s_st, s_fr)
    c = stream_mod.continuity(out)
    out_full = out[out["complete"] == 1]
    sps = d.params.interp * fs / 16640.3
    if serial_locked:
        # the reference chain itself decodes this stream: the stitched result must be the same list of minor frames
        assert serial_full.size >= min_frames and stream_mod.continuity(serial)["counter_breaks"] == 0
        assert c["counter_breaks"] == 0, c
        assert out_full.size == serial_full.size
        assert np.array_equal(out_full["bytes"], serial_full["bytes"])
        assert np.abs(out_full["sample_index"].astype(np.int64) - serial_full["sample_index"].astype(np.int64)).max() <= sps
    else:
        # the reference's own acquisition sweep never latches on this stream (2 Msps: -1 dB per-sample SNR), so the serial
        # chain decodes little or nothing; the segments behind the first still do.  Segment 0 IS the serial chain.
        own = out[out["sample_index"] >= (segment + plan.lead) * d.params.interp]
        own_full = own[own["complete"] == 1]
        assert stream_mod.continuity(own)["counter_breaks"] == 0
        assert own_full.size >= int((total - segment - plan.lead) / fs * 10) - 1
        assert np.mean([check_parity(f["bytes"]) for f in own_full]) > 0.98
    # segment 0 is the serial chain itself: the frames it owns are bit-identical including positions
    n0 = int((out["sample_index"] < (segment + plan.lead) * d.params.interp).sum())
    ns = int((serial["sample_index"] < (segment + plan.lead) * d.params.interp).sum())
    assert n0 == ns and np.array_equal(out[:n0], serial[:n0])
    assert n0 >= 3 or not serial_locked
    # a second "rank": only the slice for segments [h, k), generated separately from the same seed
    h = k // 2
    sd2 = stream_mod.StreamDemod("f32", p, plan, h, k - h)
    d_slice = torch.empty(sd2.n_slice * 2, dtype=torch.float32, device="cuda")
    assert L.pdt_synth_poes_stream_device(d_slice.data_ptr(), 0, sd2.start, sd2.n_slice, total, float(fs), 99, cs) == 0
    torch.cuda.synchronize()
    assert torch.equal(d_slice, d_iq[sd2.start * 2: (sd2.start + sd2.n_slice) * 2])
    sd2.run_device(d_slice.data_ptr(), stream=cs)
    st2, fr2 = sd2.fetch(cs)
    part2 = sd2.stitch_local(st2, fr2)
    part1 = stream_mod.stitch("f32", plan, 0, h, s_st[:h], s_fr[:h])
    assert np.array_equal(np.concatenate([part1, part2]), out)


def test_config4_stream_2msps_locked_matches_serial_and_oracle(torch_cuda, oracle32):
    """BASELINE configs[4]'s rate where the reference chain DOES lock (the synthetic device stream of the test above is too
    noisy for the reference's own acquisition sweep at 2 Msps): a 12 M-sample 2 Msps recording — the very one
    tests/test_prelock_model.py runs through the CPU model of the stream mode — demodulated (a) serially, compared byte for
    byte with oracle.chain(force_min_L1), and (b) as six pre-locked segments stitched by ownership windows: the same list of
    minor frames, every byte, positions within one symbol."""
    torch = torch_cuda
    stream_mod = importlib.import_module("project-desert-tortoise_b200.stream")
    fs, total, segment = 2_000_000, 12_000_000, 2_000_000
    pcm, _ = make_poes_capture(total, fs, 9, esn0_db=24.0, doppler_hz=1200.0, drift_hz_s=-150.0, amplitude=0.3)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, force_min_L1=True)
    assert want["locked"] and want["total_frames"] >= 55
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    p.force_min_interp1 = 1
    cs = torch.cuda.current_stream().cuda_stream
    d_iq = torch.from_numpy(iq).cuda()
    d = pdt.Demod("f32", p, 1, total, int(total / fs * 10) + 8)
    d.demod_device(d_iq.data_ptr(), 1, total, stream=cs)
    st, fr = d.fetch(1, cs)
    assert (st[0]["n_symbols"], st[0]["n_bits"], st[0]["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert st[0]["locked"] == 1 and st[0]["lock_sample"] == want["lock_sample"]
    _frames_text_equal_bytes(d.format_frames(fr[0], int(st[0]["n_frames"])), want["text"])
    serial = fr[0][: int(st[0]["n_frames"])]
    serial_full = serial[serial["complete"] == 1]
    plan = stream_mod.make_plan("f32", d.params, total, segment)
    sd = stream_mod.StreamDemod("f32", p, plan, 0, plan.n_segments)
    sd.run_device(d_iq.data_ptr(), stream=cs)
    s_st, s_fr = sd.fetch(cs)
    out = sd.stitch_local(s_st, s_fr)
    out_full = out[out["complete"] == 1]
    assert stream_mod.continuity(out)["counter_breaks"] == 0
    assert out_full.size == serial_full.size and np.array_equal(out_full["bytes"], serial_full["bytes"])
    assert np.abs(out_full["sample_index"].astype(np.int64) - serial_full["sample_index"].astype(np.int64)).max() <= fs / 16640.3


def test_config4_serial_2msps_locked_bit_exact_all_stages(torch_cuda, oracle32):
    """BASELINE configs[4]'s rate with oracle EQUALITY (not counts): 10 M samples @ 2 Msps that the serial chain does lock
    on, through the tiled engine with the declared L = max(1, …) deviation, against oracle.chain(force_min_L1) — the mode
    tests/test_oracle_vs_ref.py pins to the reference rebuilt with that one line patched.  PLL phase / output, FIR, AGC,
    symbols, Gardner errors and pick indices, bits and frame bytes are all bit-exact."""
    fs, n = 2_000_000, 10_000_000
    pcm, _ = make_poes_capture(n, fs, 7, esn0_db=24.0, doppler_hz=900.0, amplitude=0.3)
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, fs, force_min_L1=True, trace=True)
    assert want["locked"] and want["total_frames"] >= 45
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, "f32", pdt.PDT_MODE_POES, fs, iq, force_l1=True, engine="tiled")
    assert d.params.interp == 1 and d.params.taps == 26
    assert (st["n_symbols"], st["n_bits"], st["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    assert st["locked"] == 1 and st["lock_sample"] == want["lock_sample"]
    ns, nb = int(st["n_symbols"]), int(st["n_bits"])
    for k_dev, k_or in (("pll_phase", "tr_phase"), ("pll_out", "tr_pll_out"), ("lpf", "tr_lpf"), ("agc", "tr_agc")):
        assert np.array_equal(tr.host(k_dev), want[k_or]), k_dev
    assert np.array_equal(tr.host("sym", ns), want["tr_sym"])
    assert np.array_equal(tr.host("gardner_err", ns), want["tr_gerr"])
    assert np.array_equal(tr.host("gardner_idx", ns).astype(np.uint64), want["tr_gidx"])
    assert np.array_equal(tr.host("bits", nb), want["tr_bits"])
    text = d.format_frames(fr, int(st["n_frames"]))
    _frames_text_equal_bytes(text, want["text"])
    full = [f for f in parse_frames_text(text) if f[2].size == 104]
    cnt = [frame_counter(f[2]) for f in full]
    assert len(full) >= 45 and all((b - a) % 320 == 1 for a, b in zip(cnt, cnt[1:]))


def test_config2_argos_batch_of_bursts(torch_cuda, oracle64):
    """BASELINE configs[2] shape: 256 synthetic 401.65 MHz ARGOS bursts (128 captures x 2) as independent double-precision
    captures in one batch (exact engine): packets and counts of every capture equal to the CPU oracle."""
    caps, n = 128, 40_000
    iq = np.zeros((caps, n, 2), np.float64)
    wants = []
    for c in range(caps):
        pcm, _ = make_argos_capture(n, 5000.0, seed=40 + c, n_bursts=2, snr_db=14.0 + (c % 12))
        x = oracle64.pcm16_to_complex(pcm)
        iq[c] = x.reshape(-1, 2)
        wants.append(oracle64.chain(x, 5000, argos=True))
    p = pdt.default_params("f64", pdt.PDT_MODE_ARGOS, 5000)
    d = pdt.Demod("f64", p, caps, n, 16)
    st, fr = d.demod_host(iq, caps)
    assert sum(w["total_frames"] for w in wants) >= caps
    for c, w in enumerate(wants):
        assert (st[c]["n_symbols"], st[c]["n_bits"], st[c]["n_frames"]) == (w["total_symbols"], w["total_bits"], w["total_frames"]), c
        _frames_text_equal_bytes(d.format_frames(fr[c], int(st[c]["n_frames"])), w["text"])


def test_async_host_api_contexts_in_rotation(torch_cuda):
    """pdt_demod_host_async + pdt_fetch with two contexts used in rotation on two streams (what bench.py's e2e leg does):
    every batch must give exactly the synchronous pdt_demod_host result, whatever overlaps with it."""
    torch = torch_cuda
    fs, n, caps = 250000, 150_000, 130                 # > 64 captures: capture groups + chunked staging are exercised
    batches = []
    for b in range(3):
        pcm = np.zeros((caps, n, 2), np.int16)
        for c in range(0, caps, 13):                   # a few real captures, the rest noise
            x, _ = make_poes_capture(n, fs, 500 + 31 * b + c, esn0_db=15.0, doppler_hz=-2500.0 + 37.0 * c, amplitude=0.2)
            pcm[c] = x.reshape(-1, 2)
        rng = np.random.default_rng(b)
        noise = (rng.standard_normal((caps, n, 2)) * 200).astype(np.int16)
        pcm = np.where(pcm.any(axis=(1, 2), keepdims=True), pcm, noise)
        batches.append(np.ascontiguousarray(pcm))
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    ref = pdt.Demod("f32", p, caps, n, 16)
    want = [ref.demod_host(x, caps, pcm16=True) for x in batches]
    ctxs = [pdt.Demod("f32", p, caps, n, 16) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    pinned = [torch.from_numpy(x).pin_memory() for x in batches]
    got = [None] * len(batches)
    pending = [None, None]
    for i, x in enumerate(pinned + pinned[:1]):        # 4 submissions: the last one re-uses a context twice in a row
        k = i % 2
        if pending[k] is not None:
            got[pending[k]] = ctxs[k].fetch(caps, streams[k].cuda_stream)
        ctxs[k].demod_host_async(x.numpy(), caps, pcm16=True, stream=streams[k].cuda_stream)
        pending[k] = i % len(batches)
    for k in range(2):
        got[pending[k]] = ctxs[k].fetch(caps, streams[k].cuda_stream)
    for (ws, wf), (gs, gf) in zip(want, got):
        assert ws["n_frames"].sum() > 20
        assert np.array_equal(ws, gs) and np.array_equal(wf, gf)


def test_frame_post_checks_on_device(torch_cuda, golden_dir):
    """Parity word 103 / counter continuity / spacecraft id computed on the device (checkParity.m, daytimeDecode.m) against
    the numpy restatement, on the reference's own recording and on a synthetic capture with a corrupted frame."""
    rate, pcm = po.read_wav_pcm16(_golden(golden_dir, "5sec_clip.wav"))
    d = pdt.Demod("f32", pdt.default_params("f32", pdt.PDT_MODE_POES, rate), 1, pcm.size // 2, 64)
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    q = d.frame_checks(1)[0]
    nf = int(st[0]["n_frames"])
    prev = None
    seen = 0
    for f in range(nf):
        b = fr[0][f]["bytes"]
        full = bool(fr[0][f]["complete"]) and fr[0][f]["n_bytes"] == 104
        assert bool(q[f]["valid"]) == full
        if not full:
            continue
        seen += 1
        assert q[f]["counter"] == frame_counter(b) and q[f]["spacecraft"] == b[2]
        assert bool(q[f]["parity_ok"]) == check_parity(b)
        assert bool(q[f]["continuous"]) == (prev is None or frame_counter(b) == (prev + 1) % 320)
        prev = frame_counter(b)
    assert seen >= 40 and q[:nf]["spacecraft"][q[:nf]["valid"] == 1].tolist() == [8] * seen      # NOAA-15, SURVEY §8c
    assert not q[nf:]["valid"].any()


def test_frame_table_overflow_same_in_both_engines(torch_cuda):
    """More frames than max_frames slots: both engines count every accepted sync word but store only the first slots —
    identical tables, including the trailing partial frame bookkeeping."""
    fs, n = 250000, 400_000
    pcm, _ = make_poes_capture(n, fs, 61, esn0_db=16.0, doppler_hz=900.0, amplitude=0.3)
    res = {}
    for eng in ("exact", "tiled"):
        p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
        p.engine = ENGINES[eng]
        d = pdt.Demod("f32", p, 1, n, 3)
        res[eng] = d.demod_host(pcm, 1, pcm16=True)
    (se, fe), (st, ft) = res["exact"], res["tiled"]
    assert se[0]["n_frames"] > 3
    for k in ("n_symbols", "n_bits", "n_frames", "lock_sample"):
        assert se[0][k] == st[0][k], k
    assert np.array_equal(fe, ft)


def test_poes_golden_synth_c2(torch_cuda, golden_dir):
    g = np.load(_golden(golden_dir, "synth_poes_c2_small.npz"))
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, int(g["fs"]))
    d = pdt.Demod("f32", p, 1, g["pcm"].size // 2, 64)
    st, fr = d.demod_host(g["pcm"], 1, pcm16=True)
    assert d.format_frames(fr[0], int(st[0]["n_frames"])) == str(g["frames_text"])


def test_batch_many_ragged_captures(torch_cuda, oracle32):
    """Several independent captures of different length / Doppler / SNR in ONE launch; empty tail capture too."""
    fs = 250000
    lens = [260000, 123457, 10000, 9999, 300001, 64]
    stride = max(lens)
    pcm = np.zeros((len(lens), stride, 2), np.int16)
    texts = []
    for c, n in enumerate(lens):
        x, _ = make_poes_capture(n, fs, 100 + c, esn0_db=9.0 + 2 * c, doppler_hz=-3000.0 + 1100 * c, amplitude=0.08 + 0.07 * c)
        pcm[c, :n] = x.reshape(-1, 2)
        texts.append(oracle32.chain(oracle32.pcm16_to_complex(x), fs))
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    d = pdt.Demod("f32", p, len(lens), stride, 64)
    st, fr = d.demod_host(pcm, len(lens), pcm16=True, n_samples=lens)
    for c, n in enumerate(lens):
        w = texts[c]
        assert (st[c]["n_samples"], st[c]["n_symbols"], st[c]["n_bits"], st[c]["n_frames"]) == \
            (n, w["total_symbols"], w["total_bits"], w["total_frames"]), c
        _frames_text_equal_bytes(d.format_frames(fr[c], int(st[c]["n_frames"])), w["text"])


def test_l0_reference_emits_nothing_and_declared_deviation(torch_cuda, oracle32):
    rng = np.random.default_rng(0)
    pcm, _ = make_poes_capture(400000, 2_000_000, 3, esn0_db=14.0, doppler_hz=900.0)
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, 2_000_000)
    d = pdt.Demod("f32", p, 1, 400000, 16)
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    assert st[0]["n_symbols"] == 0 and st[0]["n_frames"] == 0          # like the reference (L = 0)
    p.force_min_interp1 = 1
    d = pdt.Demod("f32", p, 1, 400000, 16)
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    want = oracle32.chain(oracle32.pcm16_to_complex(pcm), 2_000_000, force_min_L1=True)
    assert (st[0]["n_symbols"], st[0]["n_bits"], st[0]["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])


@pytest.mark.parametrize("engine", ["exact", "tiled"])
@pytest.mark.parametrize("frame_len,start_bit", [(103, 3), (40, 5), (8, 0)])
def test_generic_bytesync_on_the_batch_api(torch_cuda, oracle32, engine, frame_len, start_bit):
    """the parameterised sync of common/ByteSync.c:16-144 (frameLength / startBit) as batch parameters, both
    engines, against its restatement (pinned to the reference file in tests/test_oracle_vs_ref.py) fed with the oracle's bits."""
    fs = 250000
    pcm, _ = make_poes_capture(600_000, fs, 71, esn0_db=15.0, doppler_hz=-1200.0, amplitude=0.25)
    want = oracle32.chain(oracle32.pcm16_to_complex(pcm), fs, trace=True)
    st_o = oracle32.new_state("bytesync")
    n_sync = oracle32.bytesync_generic(st_o, want["tr_bits"], b"1110110111100010000", frame_len, start_bit)
    rows = [(r[1], bytes(r[2])) for r in parse_frames_text(oracle32.bytesync_text(st_o))]
    assert n_sync >= 20
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    p.engine = ENGINES[engine]
    p.sync_generic, p.sync_frame_len, p.sync_start_bit = 1, frame_len, start_bit
    d = pdt.Demod("f32", p, 1, pcm.size // 2, 256)
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    assert int(st[0]["n_frames"]) == n_sync and int(st[0]["n_bits"]) == want["total_bits"]
    got = [(bool(f["inverse"]), bytes(f["bytes"][: f["n_bytes"]])) for f in fr[0][:n_sync]]
    assert got == rows
    assert all(len(b) == frame_len + 1 for _, b in got[:-1])


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_mm_clock_recovery_on_the_batch_api(torch_cuda, prec):
    """MMClockRecovery (common/MMClockRecovery.c:5-84 — the call both drivers keep commented out,
    POESTIPdemod/main.c:435 / ARGOSdemod/main.c:277) selected on the batch API, against the oracle's stages composed in the
    driver's order: AGC (+squelch) output of the oracle chain -> mm per chunk -> Manchester -> ByteSync."""
    o = po.Oracle(prec)
    if prec == "f32":
        fs, mode, argos, thr = 250000, pdt.PDT_MODE_POES, False, 1.0
        pcm, _ = make_poes_capture(400_000, fs, 72, esn0_db=16.0, doppler_hz=600.0, amplitude=0.25)
        iq = o.pcm16_to_complex(pcm)
        baud, chunk = 8320 * 2 + 0.3, 10000
    else:
        fs, mode, argos, thr = 5000, pdt.PDT_MODE_ARGOS, True, 0.5
        pcm, _ = make_argos_capture(60000, 5000.0, seed=9, n_bursts=3, snr_db=20.0)
        iq = o.pcm16_to_complex(pcm)
        baud, chunk = 800.0, 2400
    want = o.chain(iq, fs, argos=argos, trace=True)
    L = max(want["L"], 1)
    z = want["tr_agc"]                                                    # AGC (ARGOS: + squelch) output, per interpolated sample
    mst, manst, bst = o.new_state("mm"), o.new_state("manchester"), o.new_state("bytesync")
    syms, n_frames = [], 0
    for base in range(0, z.size, chunk * L):
        m = min(chunk * L, z.size - base)
        buf = np.zeros(m + 16, o.dt)
        buf[:m] = z[base: base + m]
        s_ = o.mm(mst, buf, m, int(np.float32(fs) * L) if prec == "f32" else int(fs), baud, 3.0, 0.15)
        syms.append(s_)
        bits = o.manchester(manst, s_, thr)
        n_frames += o.bytesync(bst, bits, kind="argos" if argos else "poes")
    syms = np.concatenate(syms)
    rows = [(r[1], bytes(r[2])) for r in parse_frames_text(o.bytesync_text(bst))]
    p = pdt.default_params(prec, mode, fs)
    p.clock_recovery = pdt.PDT_CLOCK_MM
    d = pdt.Demod(prec, p, 1, iq.size // 2, 64)
    assert d.engine == pdt.PDT_ENGINE_EXACT
    st, fr = d.demod_host(iq, 1)
    assert int(st[0]["n_symbols"]) == syms.size and int(st[0]["n_frames"]) == n_frames
    got = [(bool(f["inverse"]), bytes(f["bytes"][: f["n_bytes"]])) for f in fr[0][: min(n_frames, 64)]]
    assert got == rows[: len(got)]
    if prec == "f32":
        assert n_frames >= 10


def test_argos_synthetic_bursts(torch_cuda, oracle64):
    pcm, info = make_argos_capture(120000, 5000.0, seed=5, n_bursts=6, snr_db=18.0)
    iq = oracle64.pcm16_to_complex(pcm)
    want = oracle64.chain(iq, 5000, argos=True)
    p = pdt.default_params("f64", pdt.PDT_MODE_ARGOS, 5000)
    d = pdt.Demod("f64", p, 1, 120000, 32)
    st, fr = d.demod_host(iq, 1)
    assert (st[0]["n_symbols"], st[0]["n_bits"], st[0]["n_frames"]) == (want["total_symbols"], want["total_bits"], want["total_frames"])
    _frames_text_equal_bytes(d.format_frames(fr[0], int(st[0]["n_frames"])), want["text"])


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("bw_acq,bw_track,thresh", [(0.3, 0.05, 0.2), (0.9, 0.4, 0.03), (0.02, 0.002, 0.1)])
def test_pll_block_runner_any_loop_bandwidth(torch_cuda, oracle32, oracle64, prec, bw_acq, bw_track, thresh):
    """The block runner (pdt_pll_pipe.cuh) uses select forms of the loop filter only while one 2π wrap per sample suffices;
    loop bandwidths far outside any sane setting take the reference-shaped while-loops.  Both paths, acquisition, latch and
    track mode, call after call with carried state, against the oracle's CarrierTrackPLL: float bit for bit."""
    o = oracle64 if prec == "f64" else oracle32
    fs = 50000
    pcm, _ = make_poes_capture(30000, fs, 21, esn0_db=16.0, doppler_hz=900.0, amplitude=0.3)
    iq = o.pcm16_to_complex(pcm)
    args = (float(fs), 4500.0, thresh, 0.002, bw_acq, bw_track)
    lg = pdt.Legacy(prec)
    lg.reset()
    st = o.new_state("pll")
    cuts = [0, 7001, 7002, 15000, 15129, 30000]              # odd call lengths: partial blocks, a one-sample call
    locked_seen = False
    for a, b in zip(cuts[:-1], cuts[1:]):
        got_o, got_l, got_avg = lg.CarrierTrackPLL(iq[2 * a: 2 * b], *args, want_lock=True)
        want_o, want_l, want_avg, _, _ = o.pll(st, iq[2 * a: 2 * b], *args, want_lock=True)
        if prec == "f32":
            assert np.array_equal(got_o, want_o) and np.array_equal(got_l, want_l) and got_avg == want_avg
        else:
            np.testing.assert_allclose(got_o, want_o, rtol=1e-11, atol=1e-14)
            np.testing.assert_allclose(got_l, want_l, rtol=1e-11, atol=1e-14)
        locked_seen |= bool((want_l > thresh).any())
    assert locked_seen                                         # the latch fired somewhere: both modes were exercised


@pytest.mark.parametrize("prec,mode,fs,taps", [("f64", pdt.PDT_MODE_ARGOS, 5000, 600), ("f32", pdt.PDT_MODE_POES, 50000, 3 * 300)])
def test_exact_engine_long_fir_history_across_chunks(torch_cuda, oracle32, oracle64, prec, mode, fs, taps):
    """the FIR history (K - 1 samples carried from chunk to chunk) may be longer than the CTA — K - 1 = 599
    (plain 600-tap filter) and 299 (900 taps over L = 3) against 128 threads.  The filtered stream of the exact engine over
    five chunks must equal the oracle's stateful filter run over the engine's own PLL output, bit for bit."""
    o = oracle64 if prec == "f64" else oracle32
    n = 12000
    if mode == pdt.PDT_MODE_ARGOS:
        pcm, _ = make_argos_capture(n, float(fs), seed=11, n_bursts=1, snr_db=20.0)
    else:
        pcm, _ = make_poes_capture(n, fs, 12, esn0_db=14.0, doppler_hz=300.0, amplitude=0.25)
    iq = o.pcm16_to_complex(pcm)
    d, st, fr, tr = _run_batch_with_traces(torch_cuda, prec, mode, fs, iq, chunk=2400, engine="exact", taps=taps)
    L = max(d.params.interp, 1)
    assert d.params.taps == taps and (taps // L) - 1 > 128
    h = d.taps()
    pll_out = tr.host("pll_out", n)
    if mode == pdt.PDT_MODE_ARGOS:
        want = o.fir(o.new_state("fir"), pll_out, h)
    else:
        want = o.fir_interp(o.new_state("fir"), np.arange(n + 1, dtype=pll_out.dtype), pll_out, h, L)[0]
    got = tr.host("lpf", n * L)
    assert np.array_equal(got, np.asarray(want).ravel()[: n * L])


# ------------------------------------------------------------------------------------------------------
# legacy ABI: the reference's own function signatures, stage by stage, against the golden stage vectors
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_legacy_stage_vectors(torch_cuda, prec, golden_dir, tmp_path):
    v = np.load(_golden(golden_dir, f"stage_vectors_{prec}.npz"))
    lg = pdt.Legacy(prec)
    lg.reset()
    dt = lg.dt
    exact = prec == "f32"
    iq = v["iq"]
    n = iq.size // 4
    sg = lg.StaticGain(iq[: 2 * n])
    assert sg == float(v["static_gain"]) if exact else abs(sg / float(v["static_gain"]) - 1) < 1e-15
    a = [float(x) for x in v["pll_args"]]
    o1, l1, a1 = lg.CarrierTrackPLL(iq[: 2 * n], *a, want_lock=True)
    o2, l2, a2 = lg.CarrierTrackPLL(iq[2 * n:], *a, want_lock=True)
    got_o, got_l = np.concatenate([o1, o2]), np.concatenate([l1, l2])
    if exact:
        assert np.array_equal(got_o, v["pll_out"]) and np.array_equal(got_l, v["pll_lock"])
        assert [a1, a2] == list(v["pll_avg"])
    else:
        np.testing.assert_allclose(got_o, v["pll_out"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(got_l, v["pll_lock"], rtol=1e-12, atol=1e-15)
    x = v["fir_x"]
    if prec == "f32":
        for L in (1, 3, 8):
            lg.reset()
            h = lg.MakeLPFIR(26 * L, 11000.0, np.float32(150000.0), L)
            assert np.array_equal(h, v[f"h_L{L}"])
            tin = np.arange(2 * n + 1, dtype=dt)
            ya, ta = lg.LowPassFilterInterp(tin[: n + 1], x[:n], h, L)
            yb, tb = lg.LowPassFilterInterp(tin[n:], x[n:], h, L)
            assert np.array_equal(np.concatenate([ya, yb]), v[f"fir_interp_L{L}"])
            assert np.array_equal(ta, np.repeat(tin[1: n + 1], L))
    lg.reset()
    h = lg.MakeLPFIR(50, 700.0, 5000.0, 1)
    assert np.array_equal(h, v["h_argos"])
    assert np.array_equal(np.concatenate([lg.LowPassFilter(x[:n], h), lg.LowPassFilter(x[n:], h)]), v["fir_plain"])
    ya = lg.NormalizingAGC(v["agc_x"][:n], 17.5, 0.0033, 0.0067)
    yb = lg.NormalizingAGC(v["agc_x"][n:], 17.5, 0.0033, 0.0067)
    assert np.array_equal(np.concatenate([ya, yb]), v["agc_y"])
    FsI, baud, rng_, kp = v["gar_args"]
    buf = np.zeros(n + 16, dt)
    syms, idxs = [], []
    for k in range(2):
        buf[:n] = v["gar_x"][k * n:(k + 1) * n]
        s, i = lg.GardenerClockRecovery(buf, n, int(FsI), float(baud), float(rng_), float(kp))
        syms.append(s)
        idxs.append(i.astype(np.int64) + k * n)
    assert np.array_equal(np.concatenate(syms), v["gar_sym"])
    assert np.array_equal(np.concatenate(idxs), v["gar_idx"])
    thr = 1.0 if prec == "f32" else 0.5
    bits = np.concatenate([lg.ManchesterDecode(v["man_sym"][:1777], thr), lg.ManchesterDecode(v["man_sym"][1777:], thr)])
    assert np.array_equal(bits, v["man_bits"])


@pytest.mark.parametrize("name", ["kat_line8", "kat_line10"])
def test_legacy_bytesync_kat(torch_cuda, golden_dir, tmp_path, name):
    kat = json.load(open(_golden(golden_dir, "bytesync_kat.json")))[name]
    lg = pdt.Legacy("f32")
    lg.reset()
    bits = np.frombuffer(kat["bits"].encode(), np.uint8)
    out = str(tmp_path / "frames.txt")
    n = 0
    for lo, hi in ((0, 1), (1, 777), (777, 778), (778, bits.size)):
        n += lg.ByteSync(bits[lo:hi], out)
    assert n == kat["frames"]
    assert open(out).read() == kat["text"]


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_legacy_agcc_and_signal_amplitude(torch_cuda, prec):
    """NormalizingAGCC (AGC.c:164-200) and FindSignalAmplitude (AGC.c:6-20) of the legacy ABI against the oracle restatement
    (itself pinned to the unmodified reference in tests/test_oracle_vs_ref.py): state carried across calls; float build
    bit-exact (including its |Re| quirk), double build within 1e-13 (hypot)."""
    o = po.Oracle(prec)
    lg = pdt.Legacy(prec)
    lg.reset()
    rng = np.random.default_rng(31)
    ast = o.new_state("agc")
    avg = np.zeros(1, o.dt)
    for n in (1, 7, 2400, 10000, 3):
        iq = (rng.standard_normal(2 * n) * rng.choice([0.05, 1.0, 4.0])).astype(o.dt)
        want, got = o.agcc(ast, iq, 2.5, 1e-3), lg.NormalizingAGCC(iq, 2.5, 1e-3)
        x = (rng.standard_normal(n) * 3.0).astype(o.dt)
        wa, ga = o.signal_amplitude(avg, x, 0.01), lg.FindSignalAmplitude(x, 0.01)
        if prec == "f32":
            assert np.array_equal(want, got) and wa == ga
        else:
            np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
            assert abs(ga / wa - 1) < 1e-14


def test_legacy_squelch_mm(torch_cuda, oracle64):
    lg = pdt.Legacy("f64")
    lg.reset()
    rng = np.random.default_rng(23)
    x = rng.standard_normal(5000)
    lock = rng.random(5000) * 0.3
    assert np.array_equal(lg.Squelch(x, lock, 0.15), oracle64.squelch(x, lock, 0.15))
    st = oracle64.new_state("mm")
    for n in (500, 2400):
        buf = np.zeros(n + 16)
        buf[:n] = np.sin(np.arange(n) * 0.5) + 0.2 * rng.standard_normal(n)
        got, _ = lg.GardenerClockRecovery(buf, n, 5000, 800.0, 3.0, 0.15, mm=True)
        assert np.array_equal(got, oracle64.mm(st, buf, n, 5000, 800.0, 3.0, 0.15))


# ------------------------------------------------------------------------------------------------------
# drop-in proof: the reference's UNMODIFIED main.c + wave.c linked against libpdt (built in the container)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("app,wav,golden", [("POES", "5sec_clip.wav", "poes_5sec_clip_frames.txt"),
                                            ("ARGOS", "argos_401650kHz.wav", "argos_packets.txt")])
def test_reference_main_linked_against_shim(torch_cuda, golden_dir, tmp_path, app, wav, golden):
    exe = os.path.join(ROOT, "build", f"demod{app}_pdt")
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not prebuilt (needs /root/reference at build time)")
    r = subprocess.run([exe, _golden(golden_dir, wav)], cwd=tmp_path, capture_output=True, text=True, errors="replace",
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    outs = [f for f in os.listdir(tmp_path) if f.startswith(("minorFrames_", "packets_"))]
    assert outs, r.stdout[-2000:]
    text = open(tmp_path / outs[0]).read()
    assert text == open(_golden(golden_dir, golden)).read()
    assert "PLL locked at" in r.stdout
