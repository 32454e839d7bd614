#!/usr/bin/env python
"""tools/timeline_stream.py — per-stream kernel timeline of one stream-mode step (pdt_set_profiling(ctx, 2)).

    python tools/timeline_stream.py [--stream 1000000000] [--fs 2000000] [--segment 0]
Prints, per internal stream (capture group 0…, 99 = slow-capture stream), every kernel with its end time since the fork,
plus segment 0's lock sample (the only segment that runs the reference's serial acquisition)."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pdt = importlib.import_module("project-desert-tortoise_b200")
sm = importlib.import_module("project-desert-tortoise_b200.stream")
ap = argparse.ArgumentParser()
ap.add_argument("--stream", type=int, default=1_000_000_000)
ap.add_argument("--fs", type=int, default=2_000_000)
ap.add_argument("--segment", type=int, default=0)
a = ap.parse_args()
L = sm._bind(pdt.load("f32"))
p = pdt.default_params("f32", pdt.PDT_MODE_POES, a.fs)
if a.fs > 300000:
    p.force_min_interp1 = 1
plan = sm.make_plan("f32", p, a.stream, a.segment or 2 * a.fs)
sd = sm.StreamDemod("f32", p, plan, 0, plan.n_segments)
d_iq = torch.empty(sd.n_slice * 2, dtype=torch.float32, device="cuda")
assert L.pdt_synth_poes_stream_device(d_iq.data_ptr(), 0, 0, sd.n_slice, a.stream, float(a.fs), 20261017, 0) == 0
for _ in range(2):
    sd.run_device(d_iq.data_ptr())
torch.cuda.synchronize()
sd.demod.set_profiling(2)
sd.run_device(d_iq.data_ptr())
torch.cuda.synchronize()
tl = sd.demod.timeline()
st, fr = sd.fetch()
print("segments", plan.n_segments, "segment", plan.segment, "lead", plan.lead, "tail", plan.tail)
print("segment 0: locked", int(st[0]["locked"]), "lock_sample", int(st[0]["lock_sample"]), "of", int(st[0]["n_samples"]))
print("speculation counters (pll re-run, agc re-run, acq restarts, pll tiles):", sd.demod.tiled_counters())
by = {}
for name, g, t in tl:
    by.setdefault(g, []).append((name, t))
for g in sorted(by):
    prev = 0.0
    print(f"stream {g}:")
    for name, t in by[g]:
        print(f"   {name:16s} end {t:9.3f} ms   (+{t - prev:8.3f})")
        prev = t
