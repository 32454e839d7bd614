#!/usr/bin/env python
"""tools/seam_diag.py — look at one seam of a stream: frames of segments [s-1, s+1] next to a single long capture over the same span."""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pdt = importlib.import_module("project-desert-tortoise_b200")
sm = importlib.import_module("project-desert-tortoise_b200.stream")
ap = argparse.ArgumentParser()
ap.add_argument("--stream", type=int, default=1_000_000_000)
ap.add_argument("--fs", type=int, default=250_000)
ap.add_argument("--seg", type=int, default=1182)
a = ap.parse_args()
L = sm._bind(pdt.load("f32"))
p = pdt.default_params("f32", pdt.PDT_MODE_POES, a.fs)
plan = sm.make_plan("f32", p, a.stream, 2 * a.fs)
first, cnt = a.seg - 2, 4
sd = sm.StreamDemod("f32", p, plan, first, cnt)
d_iq = torch.empty(sd.n_slice * 2, dtype=torch.float32, device="cuda")
assert L.pdt_synth_poes_stream_device(d_iq.data_ptr(), 0, sd.start, sd.n_slice, a.stream, float(a.fs), 20261017, 0) == 0
sd.run_device(d_iq.data_ptr())
st, fr = sd.fetch()
cnt_of = lambda f: ((int(f["bytes"][4]) & 1) << 8) | int(f["bytes"][5])
for i in range(cnt):
    s = first + i
    lo, hi = s * plan.segment + plan.lead, (s + 1) * plan.segment + plan.lead
    print(f"segment {s}: owns [{lo}, {hi}) n_frames {st[i]['n_frames']} lock_freq {st[i]['lock_freq_hz']:.1f}")
    for f in fr[i][: int(st[i]["n_frames"])]:
        g = s * plan.segment + int(f["sample_index"])
        own = lo <= g < hi
        if abs(g - lo) < 0.45 * a.fs or abs(g - hi) < 0.25 * a.fs:
            print(f"    g {g} ({'own' if own else '   '}) counter {cnt_of(f)} complete {f['complete']} inv {f['inverse']} bytes {bytes(f['bytes'][:8]).hex()}")
# one long pre-locked capture over the same span
import ctypes as C
n = sd.n_slice
d = pdt.Demod("f32", p, 1, n, int(n / a.fs * 10) + 8)
lens = np.array([n], np.uint64)
assert L.pdt_demod_segments_device(d.ctx, d_iq.data_ptr(), 0, 1, n, pdt._p(lens), 0, 0) == 0
s1, f1 = d.fetch(1)
print("single capture over the span: frames", int(s1[0]["n_frames"]))
cs = [cnt_of(f) for f in f1[0][: int(s1[0]["n_frames"])] if f["complete"]]
steps = [(b - a_) % 320 for a_, b in zip(cs, cs[1:])]
print("   counter steps != 1:", [(i, cs[i], cs[i + 1]) for i, x in enumerate(steps) if x != 1])
for f in f1[0][: int(s1[0]["n_frames"])]:
    g = sd.start + int(f["sample_index"])
    lo = a.seg * plan.segment + plan.lead
    if abs(g - lo) < 0.45 * a.fs:
        print(f"    g {g} counter {cnt_of(f)} bytes {bytes(f['bytes'][:8]).hex()}")
