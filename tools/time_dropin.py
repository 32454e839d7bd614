#!/usr/bin/env python
"""Time the drop-in: the reference's UNMODIFIED main.c + wave.c linked against libpdt_f32.so (build/demodPOES_pdt) next to
the unmodified reference built from the same sources (oracle/_ref/demodPOES_ref), on one synthetic 250 ksps WAV, at the
reference's default chunk (-c 10000) and at -c 1000000.  Wall clock of the whole process (what a user of the CLI sees:
CUDA context creation and file reading included), best of `--reps`; output files must be byte-identical.

    python tools/time_dropin.py [--samples 10000000] [--out gpurun_out/dropin.json]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def run(exe, wav, chunk, cwd):
    for f in os.listdir(cwd):
        if f.startswith("minorFrames_"):
            os.remove(os.path.join(cwd, f))
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-c", str(chunk), wav], cwd=cwd, capture_output=True, text=True, errors="replace")
    dt = time.perf_counter() - t0
    outs = [f for f in os.listdir(cwd) if f.startswith("minorFrames_")]
    text = open(os.path.join(cwd, outs[0])).read() if outs else ""
    return dt, r.returncode, text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=10_000_000)
    ap.add_argument("--fs", type=int, default=250000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunks", default="10000,1000000")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from tests.synth_ref import make_poes_capture
    from tests.golden.make_golden import write_wav
    ours = os.path.join(ROOT, "build", "demodPOES_pdt")
    ref = os.path.join(ROOT, "oracle", "_ref", "demodPOES_ref")
    res = {"samples": a.samples, "fs": a.fs, "cases": []}
    with tempfile.TemporaryDirectory() as d:
        pcm, _ = make_poes_capture(a.samples, a.fs, 4242, esn0_db=14.0, doppler_hz=-1500.0, amplitude=0.25)
        wav = os.path.join(d, "s.wav")
        write_wav(wav, a.fs, pcm)
        for chunk in [int(c) for c in a.chunks.split(",")]:
            case = {"chunk": chunk}
            texts = {}
            for name, exe in (("reference_cpu", ref), ("dropin_gpu", ours)):
                best = None
                for _ in range(a.reps):
                    dt, rc, text = run(exe, wav, chunk, d)
                    assert rc == 0, (name, rc)
                    best = dt if best is None else min(best, dt)
                texts[name] = text
                case[name] = {"seconds": round(best, 4), "Msamples_per_s": round(a.samples / best / 1e6, 2),
                              "frames": text.count("\n")}
            case["identical_output"] = texts["reference_cpu"] == texts["dropin_gpu"]
            case["speedup"] = round(case["reference_cpu"]["seconds"] / case["dropin_gpu"]["seconds"], 3)
            res["cases"].append(case)
            print(json.dumps(case), flush=True)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
