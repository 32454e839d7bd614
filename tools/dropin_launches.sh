#!/bin/bash
# Per-kernel time of the drop-in CLI (build/demodPOES_pdt) on a 1 M-sample synthetic WAV at the reference's default chunk:
# which legacy stage kernel the CLI's wall time goes to.  ncu launch list (serialised, cold-cache: shares, not absolutes).
set -e
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/dropin_launches.csv}
python - <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from tests.synth_ref import make_poes_capture
from tests.golden.make_golden import write_wav
pcm, _ = make_poes_capture(1_000_000, 250000, 4242, esn0_db=14.0, doppler_hz=-1500.0, amplitude=0.25)
write_wav("/tmp/dropin_s.wav", 250000, pcm)
PY
cd /tmp && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file /tmp/dl.csv "$OLDPWD/build/demodPOES_pdt" -c 10000 /tmp/dropin_s.wav > /dev/null 2>&1
cd "$OLDPWD"
python - "$OUT" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open("/tmp/dl.csv") if l.startswith('"'))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    k = r[ki].split("(")[0]; t[k][0] += 1; t[k][1] += v
tot = sum(v[1] for v in t.values())
with open(sys.argv[1], "w") as f:
    f.write("kernel,launches,total_us,us_per_launch,share\n")
    for k, (n, us) in sorted(t.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{us:.1f},{us/n:.2f},{us/tot:.3f}\n")
print(open(sys.argv[1]).read())
PY
