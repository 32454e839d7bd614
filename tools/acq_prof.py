#!/usr/bin/env python
"""Cycle accounting of the packed acquisition's second pass (pdt_debug_acq_prof), one batch alone."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pdt = importlib.import_module("project-desert-tortoise_b200")
L = pdt.load("f32")
L.pdt_debug_acq_prof.argtypes = [C.POINTER(C.c_uint64 * 8), C.c_int]
C_, n, FS = 1024, 1_000_000, 250000
d_iq = torch.empty(C_ * n * 2, dtype=torch.float32, device="cuda")
assert L.pdt_synth_poes_device(d_iq.data_ptr(), 0, C_, n, n, float(FS), 20261017, 0) == 0
d = pdt.Demod("f32", pdt.default_params("f32", pdt.PDT_MODE_POES, FS), C_, n, 48)
out = (C.c_uint64 * 8)()
d.demod_device(d_iq.data_ptr(), C_, n); torch.cuda.synchronize()
L.pdt_debug_acq_prof(C.byref(out), 1)
d.demod_device(d_iq.data_ptr(), C_, n); torch.cuda.synchronize()
L.pdt_debug_acq_prof(C.byref(out), 1)
v = [int(x) for x in out]
steps = max(v[0], 1)
names = ["steps (sum over CTAs)", "core warp phase A", "core warp barrier wait", "EMA warp phase A", "helper warp 3 phase A", "helper warp 3 barrier wait", "phase C control warp (incl. barrier wait)", "phase C warp 0"]
print(names[0], v[0])
for k in range(1, 8):
    print(f"{names[k]:28s} {v[k] / steps:9.0f} cycles per step")
