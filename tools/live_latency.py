#!/usr/bin/env python
"""Latency of one live-mode push (pdt_live_push_host: H2D + one chunk through the serial chain + D2H of stats and frames)
for 1 … 1024 simultaneous streams, POES at 50 ksps (the reference's sound-card rates are 32–48 kHz) with its default chunk."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
pdt = importlib.import_module("project-desert-tortoise_b200")
from tests.synth_ref import make_poes_capture
fs, chunk, pushes = 50000, 10000, 12
pcm, _ = make_poes_capture(chunk * pushes, fs, 5, esn0_db=15.0, doppler_hz=800.0, amplitude=0.25)
iq = (pcm.astype(np.float32) / np.float32(32768.0)).reshape(pushes, 2 * chunk)
out = []
for n_streams in (1, 16, 256, 1024):
    p = pdt.default_params("f32", pdt.PDT_MODE_POES, fs)
    live = pdt.Live("f32", p, n_streams, chunk, 16)
    ts = []
    frames = 0
    for k in range(pushes):
        block = np.ascontiguousarray(np.broadcast_to(iq[k], (n_streams, 2 * chunk)))
        t0 = time.perf_counter()
        done = live.push(block)
        ts.append(time.perf_counter() - t0)
        frames += len(done[0])
    live.close()
    ms = float(np.median(ts[2:]) * 1e3)
    out.append({"streams": n_streams, "chunk": chunk, "fs": fs, "ms_per_push_median": round(ms, 3),
                "chunk_duration_ms": 1e3 * chunk / fs, "realtime_factor_per_stream": round(1e3 * chunk / fs / ms, 1),
                "aggregate_Msamples_per_s": round(n_streams * chunk / ms / 1e3, 2), "frames_stream0": frames})
    print(json.dumps(out[-1]), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
