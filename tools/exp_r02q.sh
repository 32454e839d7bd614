mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
B="python bench.py --no-e2e --no-cpu --no-single --steps 20"
run() { name=$1; shift; env $ENVV timeout 300 $B "$@" > gpurun_out/r02q_$name.json 2> gpurun_out/r02q_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"], d["check"]["speculation"], [(k["kernel"],round(k["ms"],2)) for k in d["kernels"][:6]])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02q_$name.err").read()[-400:])
PY
}
ENVV="X=1" run packed
ENVV="PDT_ACQ_PACKED=0" run percta
ENVV="X=1" run packed_g3_if4 --groups 3 --inflight 4
ENVV="X=1" run packed_if12 --groups 1 --inflight 12
timeout 200 python tools/timeline_inflight.py --inflight 1 > gpurun_out/r02q_tl1.txt 2>&1; grep -A16 "stream 0" gpurun_out/r02q_tl1.txt | head -18; grep -A14 "stream 99" gpurun_out/r02q_tl1.txt
