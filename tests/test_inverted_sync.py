"""The inverted-sync false accept at a frame boundary (POESTIPdemod/ByteSync.c:120-133).

When the last 16 payload bits of a minor frame are 0001 0010 0001 1101, they spell — together with the first three bits of
the next frame's ED — the INVERTED 19-bit sync word.  The reference is out of frame at that moment and accepts it, 16 bits
early: the next minor frame comes out inverted and shifted (its row carries the 'i' mark) and the frame counter breaks twice.
This is reference behaviour (expected once per 2^16 frames with random payload), not a stitching artefact: the tests pin it on
the unmodified reference, on the oracle restatement, and (GPU) on both engines and on the stream mode around a seam.
"""
import importlib
import os

import numpy as np
import pytest

import pyoracle as po
from tests.synth_ref import frame_counter, make_poes_capture, parse_frames_text

FS = 250000


def crafted_capture(n=1_500_000, seed=12, which=(9,)):
    def hook(frames):
        for k in which:
            frames[k, 102], frames[k, 103] = 0x12, 0x1D
    return make_poes_capture(n, FS, seed, esn0_db=16.0, doppler_hz=800.0, amplitude=0.25, frames_hook=hook)


def _breaks(text):
    full = [f for f in parse_frames_text(text) if f[2].size == 104]
    cnt = [frame_counter(f[2]) for f in full if not f[1]]
    return sum((b - a) % 320 != 1 for a, b in zip(cnt, cnt[1:])), sum(1 for f in full if f[1])


def test_oracle_and_reference_accept_the_inverted_sync(tmp_path, oracle32):
    pcm, _ = crafted_capture()
    res = oracle32.chain(oracle32.pcm16_to_complex(pcm), FS)
    rows = parse_frames_text(res["text"])
    inv = [i for i, r in enumerate(rows) if r[1]]
    assert len(inv) == 1                                            # exactly one row carries the 'i' mark
    # the row in front of it ends with the crafted bytes; one minor frame is lost -> the counter of the upright rows jumps by 2
    assert rows[inv[0] - 1][2][102:104].tolist() == [0x12, 0x1D]
    up = [frame_counter(r[2]) for r in rows if r[2].size == 104 and not r[1]]
    steps = [(b - a) % 320 for a, b in zip(up, up[1:])]
    assert steps.count(2) == 1 and steps.count(1) == len(steps) - 1
    if po.ref_available("f32"):
        from tests.golden.make_golden import write_wav
        wav = str(tmp_path / "inv.wav")
        write_wav(wav, FS, pcm)
        _, txt = po.run_ref_cli("POES", wav)
        assert txt == res["text"]                                   # the unmodified reference prints the very same file


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["exact", "tiled"])
def test_gpu_engines_reproduce_the_inverted_sync(oracle32, engine):
    pdt = importlib.import_module("project-desert-tortoise_b200")
    pcm, _ = crafted_capture()
    want = oracle32.chain(oracle32.pcm16_to_complex(pcm), FS)
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, FS)
    p.engine = {"exact": pdt.PDT_ENGINE_EXACT, "tiled": pdt.PDT_ENGINE_TILED}[engine]
    d = pdt.Demod("f32", p, 1, pcm.size // 2, 80)
    st, fr = d.demod_host(pcm, 1, pcm16=True)
    assert d.format_frames(fr[0], int(st[0]["n_frames"])) == want["text"]
    assert int(fr[0]["inverse"][: int(st[0]["n_frames"])].sum()) == 1


@pytest.mark.gpu
def test_stream_mode_seam_next_to_an_inverted_sync_equals_serial(oracle32):
    """The crafted frame sits right in front of a segment seam: the stitched stream must show exactly what the serial capture
    shows (same rows, same 'i' row, the same two counter breaks) — the breaks are the reference's, not the stitch's."""
    import torch
    pdt = importlib.import_module("project-desert-tortoise_b200")
    sm = importlib.import_module("project-desert-tortoise_b200.stream")
    n, segment = 3_000_000, 500_000
    lead = int(0.3 * FS)
    # frames are 25 000 samples long; seams own from s*segment + lead: put crafted frames just before the seams of segments 2 and 4
    pcm, info = make_poes_capture(n, FS, 12, esn0_db=16.0, doppler_hz=800.0, amplitude=0.25)
    spf = FS / 16640.3 * 2 * 8 * 104
    first_start = (2 * 8 * 104 - info["bit_start"]) % (2 * 8 * 104) * (FS / 16640.3 * 2)      # sample where frame 1 of the table starts
    ks = [int((s * segment + lead - first_start) // spf) for s in (2, 4)]                          # frame whose END is nearest in front of the seam
    pcm, _ = crafted_capture(n, 12, which=tuple(k + 1 for k in ks))
    iq = oracle32.pcm16_to_complex(pcm)
    want = oracle32.chain(iq, FS)
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, FS)
    d_iq = torch.from_numpy(iq).cuda()
    cs = torch.cuda.current_stream().cuda_stream
    d = pdt.Demod("f32", p, 1, n, 160)
    d.demod_device(d_iq.data_ptr(), 1, n, stream=cs)
    st, fr = d.fetch(1, cs)
    assert d.format_frames(fr[0], int(st[0]["n_frames"])) == want["text"]
    serial = fr[0][: int(st[0]["n_frames"])]
    assert int(serial["inverse"].sum()) == 2
    plan = sm.make_plan("f32", d.params, n, segment)
    sd = sm.StreamDemod("f32", p, plan, 0, plan.n_segments)
    sd.run_device(d_iq.data_ptr(), stream=cs)
    s_st, s_fr = sd.fetch(cs)
    out = sd.stitch_local(s_st, s_fr)
    a, b = out[out["complete"] == 1], serial[serial["complete"] == 1]
    assert a.size == b.size and np.array_equal(a["bytes"], b["bytes"]) and np.array_equal(a["inverse"], b["inverse"])
    assert sm.continuity(out)["counter_breaks"] == sm.continuity(serial)["counter_breaks"] >= 2
