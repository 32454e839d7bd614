#!/usr/bin/env python
"""Cycle accounting of the exact engine (pdt_debug_chain_prof): where thread 0 of a capture's CTA spends its time.
ARGOS double batch (bench.py --mode argos shape) and a POES float batch at 50 ksps (L = 3), one batch alone each.

    python tools/chain_prof.py [out.json]
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pdt = importlib.import_module("project-desert-tortoise_b200")
NAMES = ["static_gain", "pll", "pll_P_alone", "pll_C||E", "pll_H||P", "pll_C_busy", "pll_E_busy", "pll_blocks", "pll_contradicted", "fir", "agc",
         "clock_bits", "capture_total", "samples", "pll_ctl", "pll_ctl_work", "pll_L_busy"]


def prof(prec, reset=1):
    L = pdt.load(prec)
    L.pdt_debug_chain_prof.argtypes = [C.POINTER(C.c_uint64 * 20), C.c_int]
    out = (C.c_uint64 * 20)()
    assert L.pdt_debug_chain_prof(C.byref(out), reset) == 0
    return [int(v) for v in out]


def report(tag, prec, ms, caps):
    v = prof(prec)
    n = max(v[13], 1)
    row = {"case": tag, "captures": caps, "kernel_ms": round(ms, 3), "samples": v[13],
           "cycles_per_sample": {NAMES[i]: round(v[i] / n, 1) for i in (0, 1, 2, 3, 4, 5, 6, 16, 14, 15, 9, 10, 11, 12)},
           "pll_blocks": v[7], "pll_contradicted_blocks": v[8], "samples_per_block": round(n / max(v[7], 1), 1)}
    print(json.dumps(row), flush=True)
    return row


def timed(d, ptr, caps, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d.demod_device(ptr, caps, n)
    torch.cuda.synchronize()
    prof(d.prec)
    a.record()
    d.demod_device(ptr, caps, n)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def main():
    rows = []
    import bench
    # ARGOS, double
    caps = 1024
    bench.FS = bench.ARGOS_FS
    host = bench.argos_captures(caps)
    d_iq = torch.from_numpy(host).cuda()
    p = pdt.default_params("f64", pdt.PDT_MODE_ARGOS, bench.ARGOS_FS)
    d = pdt.Demod("f64", p, caps, bench.ARGOS_N, 16)
    rows.append(report("argos_f64_1024x40k", "f64", timed(d, d_iq.data_ptr(), caps, bench.ARGOS_N), caps))
    del d, d_iq
    # POES float, exact engine, 50 ksps (L = 3)
    from tests.synth_ref import make_poes_capture
    caps, n = 256, 250_000
    base = []
    for c in range(8):
        pcm, _ = make_poes_capture(n, 50000, 100 + c, esn0_db=12.0, doppler_hz=-800.0 + 200 * c, amplitude=0.25)
        base.append((np.ascontiguousarray(pcm, np.int16) / np.float32(32768)).astype(np.float32).view(np.complex64))   # wave.c:151-156
    host = np.stack([base[c % 8] for c in range(caps)]).view(np.float32).reshape(caps, 2 * n)
    d_iq = torch.from_numpy(host).cuda()
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, 50000)
    p.engine = pdt.PDT_ENGINE_EXACT
    d = pdt.Demod("f32", p, caps, n, 64)
    rows.append(report("poes_f32_exact_256x250k_50ksps", "f32", timed(d, d_iq.data_ptr(), caps, n), caps))
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
