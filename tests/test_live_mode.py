"""Live mode (include/pdt.h pdt_live_*): the reference's sound-card loop (POESTIPdemodPortAudio/main.c:324-401,
ARGOSdemodPortAudio/main.c:290-329) — one chunk per iteration, stage state carried in statics — as pushes with carried state.
A sequence of pushes must give exactly what the reference chain gives when it is fed the same sequence of chunk lengths."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import pyoracle as po
from tests.synth_ref import make_argos_capture, make_poes_capture, parse_frames_text

pdt = importlib.import_module("project-desert-tortoise_b200")
pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_pushes(o, iq, fs, cuts, argos=False, chunk=None):
    """oracle chain fed push by push (pdto_chain_feed cuts a push longer than `chunk` like the kernel does)."""
    c = o.lib.pdto_chain_new(int(argos), float(fs), chunk or (2400 if argos else 10000), 0)
    try:
        for a, b in zip(cuts, cuts[1:]):
            piece = np.ascontiguousarray(iq[2 * a: 2 * b])
            o.lib.pdto_chain_feed(c, piece.ctypes.data, b - a)
        ln = C.c_size_t(0)
        text = C.string_at(o.lib.pdto_chain_text(c, C.byref(ln)), ln.value).decode()
        tot = o._chain_totals(c)
    finally:
        o.lib.pdto_chain_free(c)
    return text, tot


def _rows(text):
    return [(r[1], bytes(r[2])) for r in parse_frames_text(text)]


def _run_live(prec, params, iq_streams, cuts, max_chunk, max_frames=16):
    n_streams = len(iq_streams)
    live = pdt.Live(prec, params, n_streams, max_chunk, max_frames)
    got = [[] for _ in range(n_streams)]
    dt = np.float32 if prec == "f32" else np.float64
    for a, b in zip(cuts, cuts[1:]):
        block = np.stack([np.asarray(x[2 * a: 2 * b], dt) for x in iq_streams])
        for s, done in enumerate(live.push(block)):
            got[s] += [(bool(f["inverse"]), bytes(f["bytes"][: f["n_bytes"]])) for f in done]
    stats = live.stats.copy()
    pend = live.pending()
    live.close()
    return got, stats, pend


def test_live_pushes_of_the_reference_chunk_equal_the_file_result():
    """5sec_clip.wav pushed 10000 samples at a time (the reference's chunk): the minor frames of the golden file, in order."""
    rate, pcm = po.read_wav_pcm16(os.path.join(GOLDEN, "5sec_clip.wav"))
    o = po.Oracle("f32")
    iq = o.pcm16_to_complex(pcm)
    n = iq.size // 2
    cuts = list(range(0, n, 10000)) + [n]
    golden = _rows(open(os.path.join(GOLDEN, "poes_5sec_clip_frames.txt")).read())
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, rate)
    got, stats, pend = _run_live("f32", p, [iq], cuts, 10000)
    full = [g for g in golden if len(g[1]) == 104]
    assert got[0] == full and len(full) >= 45
    assert pend[0] == len(golden) - len(full)                                  # the trailing partial frame stays pending
    assert int(stats[0]["n_samples"]) == n and int(stats[0]["n_frames"]) == len(golden)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_live_ragged_pushes_equal_the_oracle_fed_the_same_way(prec):
    """Ragged push lengths (1 sample, odd sizes, longer than a chunk), three streams at once with different recordings."""
    o = po.Oracle(prec)
    rng = np.random.default_rng(8)
    if prec == "f32":
        fs, mode, argos, n, chunk = 250000, pdt.PDT_MODE_POES, False, 300_000, 10000
        iqs = [o.pcm16_to_complex(make_poes_capture(n, fs, 90 + k, esn0_db=15.0 + k, doppler_hz=-900.0 + 700 * k, amplitude=0.2)[0]) for k in range(3)]
    else:
        fs, mode, argos, n, chunk = 5000, pdt.PDT_MODE_ARGOS, True, 60_000, 2400
        iqs = [o.pcm16_to_complex(make_argos_capture(n, 5000.0, seed=20 + k, n_bursts=3, snr_db=18.0 + k)[0]) for k in range(3)]
    cuts = [0, 1, 778]
    while cuts[-1] < n:
        cuts.append(min(n, cuts[-1] + int(rng.choice([chunk, chunk, chunk // 3 + 1, 2 * chunk + 77, 999]))))
    max_chunk = max(b - a for a, b in zip(cuts, cuts[1:]))
    p = pdt.default_params(prec, mode, fs)
    got, stats, pend = _run_live(prec, p, iqs, cuts, max_chunk, max_frames=32)
    total = 0
    for s, iq in enumerate(iqs):
        text, tot = _oracle_pushes(o, iq, fs, cuts, argos=argos)
        rows = _rows(text)
        full_len = 104 if prec == "f32" else 7
        full = [r for r in rows if len(r[1]) == full_len]
        assert got[s] == full, s
        assert (int(stats[s]["n_symbols"]), int(stats[s]["n_bits"]), int(stats[s]["n_frames"])) == \
            (tot["total_symbols"], tot["total_bits"], tot["total_frames"]), s
        total += len(full)
    assert total >= (20 if prec == "f32" else 3)


def test_live_restart_gives_the_same_result_again():
    fs, n = 250000, 120_000
    o = po.Oracle("f32")
    iq = o.pcm16_to_complex(make_poes_capture(n, fs, 95, esn0_db=16.0, doppler_hz=500.0, amplitude=0.25)[0])
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    live = pdt.Live("f32", p, 1, 10000, 16)
    runs = []
    for _ in range(2):
        pdt._check(live.d.L, live.d.L.pdt_live_begin(live.d.ctx))
        live.reported = [0]
        rows = []
        for a in range(0, n, 10000):
            rows += [bytes(f["bytes"]) for f in live.push(iq[2 * a: 2 * (a + 10000)].reshape(1, -1))[0]]
        runs.append((rows, live.stats.copy()))
    live.close()
    assert runs[0][0] == runs[1][0] and len(runs[0][0]) >= 3 and np.array_equal(runs[0][1], runs[1][1])


def test_live_c_example_prints_the_golden_frames(tmp_path):
    """examples/live_stdin.c (plain C99 against include/pdt.h): 5sec_clip.wav's PCM piped through it in blocks of 10000."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "project-desert-tortoise_b200")
    exe = tmp_path / "live_stdin"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I" + os.path.join(root, "include"), "-o", str(exe),
                    os.path.join(root, "examples", "live_stdin.c"), "-L" + pkg, "-lpdt_f32", "-Wl,-rpath," + pkg, "-lm"], check=True)
    rate, pcm = po.read_wav_pcm16(os.path.join(GOLDEN, "5sec_clip.wav"))
    r = subprocess.run([str(exe), str(rate), "10000"], input=np.ascontiguousarray(pcm, np.int16).tobytes(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    golden = _rows(open(os.path.join(GOLDEN, "poes_5sec_clip_frames.txt")).read())
    got = _rows(r.stdout.decode())
    full = [g for g in golden if len(g[1]) == 104]                            # the trailing partial frame is never completed
    assert len(full) >= 45 and got == full

