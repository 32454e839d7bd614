#!/usr/bin/env python
"""tools/timeline.py — per-stream kernel timeline of one production-mode batch (pdt_set_profiling(ctx, 2)).

    python tools/timeline.py [--captures 1024] [--samples 1000000] [--pcm16]
Prints, per stream (capture group 0…5, 99 = slow-capture stream), every kernel with its end time since the fork."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pdt = importlib.import_module("project-desert-tortoise_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--captures", type=int, default=1024)
ap.add_argument("--samples", type=int, default=1_000_000)
ap.add_argument("--pcm16", action="store_true")
a = ap.parse_args()
L = pdt.load("f32")
C_, n, FS = a.captures, a.samples, 250000
d_iq = torch.empty(C_ * n * 2, dtype=torch.int16 if a.pcm16 else torch.float32, device="cuda")
assert L.pdt_synth_poes_device(d_iq.data_ptr(), int(a.pcm16), C_, n, n, float(FS), 20261017, 0) == 0
d = pdt.Demod("f32", pdt.default_params("f32", pdt.PDT_MODE_POES, FS), C_, n, int(n / FS * 10) + 8)
for _ in range(3):
    d.demod_device(d_iq.data_ptr(), C_, n, pcm16=a.pcm16)
torch.cuda.synchronize()
import ctypes
prof = (ctypes.c_uint64 * 8)()
L.pdt_debug_acq_prof(prof, 1)
d.set_profiling(2)
d.demod_device(d_iq.data_ptr(), C_, n, pcm16=a.pcm16)
torch.cuda.synchronize()
L.pdt_debug_acq_prof(prof, 0)
p = list(prof)
if p[0]:
    print('acq pass-1 pipeline: steps %d epochs %d cycles/step %.0f core busy %.0f ema busy %.0f helper busy %.0f decision phase %.0f' % (p[0], p[6], p[1]/p[0], p[2]/p[0], p[3]/p[0], p[4]/p[0], p[5]/p[0]))
tl = d.timeline()
print('speculation counters (pll re-run, agc re-run, acq restarts, pll tiles):', d.tiled_counters())
by = {}
for name, g, t in tl:
    by.setdefault(g, []).append((name, t))
for g in sorted(by):
    prev = 0.0
    print(f"stream {g}:")
    for name, t in by[g]:
        print(f"   {name:16s} end {t:9.3f} ms   (+{t - prev:8.3f})")
        prev = t
