#!/usr/bin/env python
"""tools/timeline_inflight.py — per-stream kernel timeline of ONE batch while `--inflight` batches share the GPU
(the steady state bench.py measures).  Prints, for capture group 0 and the slow-capture stream, every kernel's duration."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
pdt = importlib.import_module("project-desert-tortoise_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--captures", type=int, default=1024)
ap.add_argument("--samples", type=int, default=1_000_000)
ap.add_argument("--inflight", type=int, default=3)
a = ap.parse_args()
L = pdt.load("f32")
C_, n, FS = a.captures, a.samples, 250000
d_iq = torch.empty(C_ * n * 2, dtype=torch.float32, device="cuda")
assert L.pdt_synth_poes_device(d_iq.data_ptr(), 0, C_, n, n, float(FS), 20261017, 0) == 0
ctxs = [pdt.Demod("f32", pdt.default_params("f32", pdt.PDT_MODE_POES, FS), C_, n, int(n / FS * 10) + 8) for _ in range(a.inflight)]
streams = [torch.cuda.Stream() for _ in range(a.inflight)]
torch.cuda.synchronize()
for c in ctxs:
    c.set_profiling(2)
for i in range(4 * a.inflight):
    k = i % a.inflight
    ctxs[k].demod_device(d_iq.data_ptr(), C_, n, stream=streams[k].cuda_stream)
torch.cuda.synchronize()
tl = ctxs[a.inflight // 2].timeline()
by = {}
for name, g, t in tl:
    by.setdefault(g, []).append((name, t))
for g in sorted(by):
    prev = by[g][0][1] if by.get(g) else 0.0
    print(f"stream {g}:")
    for name, t in by.get(g, []):
        print(f"   {name:16s} end {t:9.3f} ms   (+{t - prev:8.3f})")
        prev = t
