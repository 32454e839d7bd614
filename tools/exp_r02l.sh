mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
B="python bench.py --no-e2e --no-cpu --no-single --steps 20"
run() { name=$1; shift; env $ENVV $B "$@" > gpurun_out/r02l_$name.json 2> gpurun_out/r02l_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02l_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02l_$name.err").read()[-300:])
PY
}
ENVV="PDT_LIB_VARIANT=_all" run all_if4
ENVV="PDT_LIB_VARIANT=_all" run all_if6 --inflight 6
ENVV="PDT_LIB_VARIANT=_all" run all_g1_if8 --inflight 8 --groups 1
ENVV="PDT_LIB_VARIANT=_all" run all_g2_if8 --inflight 8 --groups 2
ENVV="PDT_LIB_VARIANT=_acqls" run acqls_if4
ENVV="PDT_LIB_VARIANT=_acqls" run acqls_g1_if8 --inflight 8 --groups 1
# ARGOS bench line (configs[2]) and its reference arm
python bench.py --mode argos --steps 10 > gpurun_out/r02l_argos.json 2> gpurun_out/r02l_argos.err; tail -c 400 gpurun_out/r02l_argos.err; cut -c1-600 gpurun_out/r02l_argos.json
python bench.py --mode argos --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_argos_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r02l_argos_ref.json
# launch list of the drop-in binary
python - <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from tests.synth_ref import make_poes_capture
from tests.golden.make_golden import write_wav
pcm, _ = make_poes_capture(1_000_000, 250000, 4242, esn0_db=14.0, doppler_hz=-1500.0, amplitude=0.25)
write_wav("/tmp/s1m.wav", 250000, pcm)
PY
cd /tmp && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $OLDPWD/gpurun_out/r02l_dropin_launches.csv $OLDPWD/build/demodPOES_pdt /tmp/s1m.wav > /dev/null 2>&1; cd $OLDPWD
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02l_dropin_launches.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
acc=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: acc[r[ki]]+=float(r[vi].replace(",","")); cnt[r[ki]]+=1
    except: pass
for k,v in acc.most_common(): print(k[:50], cnt[k], round(v/1e3,1), "us total", round(v/cnt[k]/1e3,2), "us each")
PY
