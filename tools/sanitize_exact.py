#!/usr/bin/env python
"""tools/sanitize_exact.py — small runs of the exact engine (block-runner PLL, staged windows, symbol queues) and of the
legacy CarrierTrackPLL for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/sanitize_exact.py
    compute-sanitizer --tool memcheck  python tools/sanitize_exact.py
POES float at 50 ksps (L = 3) with ragged lengths, ARGOS double, live pushes of odd sizes, legacy PLL calls of odd sizes."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.synth_ref import make_argos_capture, make_poes_capture
pdt = importlib.import_module("project-desert-tortoise_b200")



def to_iq(pcm, dt):                       # int16 / 32768 like wave.c:151-156 (exact in both precisions)
    return (np.ascontiguousarray(pcm, np.int16).astype(np.float64) / 32768.0).astype(dt)

# POES float, exact engine, L = 3, two ragged captures
n = 60_000
iq = np.stack([to_iq(make_poes_capture(n, 50000, 5 + c, esn0_db=14.0, doppler_hz=500.0 * c, amplitude=0.25)[0], np.float32) for c in range(2)])
p = pdt.default_params("f32", pdt.PDT_MODE_POES, 50000)
p.engine = pdt.PDT_ENGINE_EXACT
d = pdt.Demod("f32", p, 2, n, 32)
d_iq = torch.from_numpy(np.ascontiguousarray(iq, np.float32)).cuda()
d.demod_device(d_iq.data_ptr(), 2, n, n_samples=np.array([n, n - 4321], np.uint64))
st, fr = d.fetch(2)
print("poes exact frames", st["n_frames"].tolist(), "symbols", st["n_symbols"].tolist())
d.close()
# ARGOS double
n = 40_000
iqa = to_iq(make_argos_capture(n, 5000.0, seed=3, n_bursts=2, snr_db=18.0)[0], np.float64)
p = pdt.default_params("f64", pdt.PDT_MODE_ARGOS, 5000)
d = pdt.Demod("f64", p, 2, n, 16)
st, fr = d.demod_host(np.stack([iqa, iqa]), 2)
print("argos frames", st["n_frames"].tolist(), "symbols", st["n_symbols"].tolist())
d.close()
# live pushes of odd sizes
p = pdt.default_params("f32", pdt.PDT_MODE_POES, 50000)
live = pdt.Live("f32", p, 2, 10000, 16)
cuts = [0, 10000, 13333, 13334, 23334, 30001]
for a, b in zip(cuts, cuts[1:]):
    live.push(np.ascontiguousarray(iq[:, 2 * a: 2 * b], np.float32))
print("live", live.stats["n_frames"].tolist(), live.stats["n_samples"].tolist())
live.close()
# legacy PLL, odd call lengths
lg = pdt.Legacy("f32")
lg.reset()
args = (50000.0, 4500.0, 0.1, 0.002, 0.02, 0.002)
for a, b in ((0, 7001), (7001, 7002), (7002, 20000)):
    lg.CarrierTrackPLL(iq[0][2 * a: 2 * b], *args, want_lock=True)
torch.cuda.synchronize()
print("done")
