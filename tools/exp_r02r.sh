mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "tiled or config or batch or async or golden" 2>&1 | tail -4
B="python bench.py --no-e2e --no-cpu --no-single --steps 20"
run() { name=$1; shift; env $ENVV timeout 300 $B "$@" > gpurun_out/r02r_$name.json 2> gpurun_out/r02r_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02r_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"], d["check"]["speculation"]["acq_restarts"], [(k["kernel"],round(k["ms"],2)) for k in d["kernels"][:6]])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02r_$name.err").read()[-400:])
PY
}
ENVV="X=1" run packed
ENVV="PDT_ACQ_PACKED=0" run percta
timeout 200 python tools/timeline_inflight.py --inflight 1 > gpurun_out/r02r_tl1.txt 2>&1; grep -A4 "stream 0" gpurun_out/r02r_tl1.txt | head -6; grep -A3 "stream 99" gpurun_out/r02r_tl1.txt
