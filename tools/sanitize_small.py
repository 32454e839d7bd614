#!/usr/bin/env python
"""tools/sanitize_small.py — small end-to-end runs of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
batch at L = 1 / 3 / 8 (tiled engine, ragged lengths, pcm16 and cf32), the exact engine, and a short stream as segments."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pdt = importlib.import_module("project-desert-tortoise_b200")
sm = importlib.import_module("project-desert-tortoise_b200.stream")
L = sm._bind(pdt.load("f32"))
for fs, n, caps, pcm in ((250000, 300_000, 5, 0), (50000, 90_000, 3, 1), (18750, 40_000, 3, 0)):
    el = torch.int16 if pcm else torch.float32
    d_iq = torch.empty(caps * n * 2, dtype=el, device="cuda")
    assert L.pdt_synth_poes_device(d_iq.data_ptr(), pcm, caps, n, n, float(fs), 7, 0) == 0
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    for engine in (pdt.PDT_ENGINE_TILED, pdt.PDT_ENGINE_EXACT):
        p.engine = engine
        p.acq_first = 40960
        d = pdt.Demod("f32", p, caps, n, 64)
        lens = np.array([n - 1237 * i for i in range(caps)], np.uint64)
        d.demod_device(d_iq.data_ptr(), caps, n, pcm16=bool(pcm), n_samples=lens)
        st, fr = d.fetch(caps)
        q = d.frame_checks(caps)
        print(fs, "engine", engine, "frames", st["n_frames"].tolist(), "locked", st["locked"].tolist())
        d.close()
fs, total = 250000, 3_000_000
p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
plan = sm.make_plan("f32", p, total, 500_000)
sd = sm.StreamDemod("f32", p, plan, 0, plan.n_segments)
d_iq = torch.empty(total * 2, dtype=torch.float32, device="cuda")
assert L.pdt_synth_poes_stream_device(d_iq.data_ptr(), 0, 0, total, total, float(fs), 99, 0) == 0
sd.run_device(d_iq.data_ptr())
st, fr = sd.fetch()
out = sd.stitch_local(st, fr)
print("stream", plan.n_segments, "segments", sm.continuity(out))
torch.cuda.synchronize()
print("done")
