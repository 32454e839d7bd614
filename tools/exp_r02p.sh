mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-single --steps 16 --warmup 16"
run() { name=$1; shift; env PDT_DEBUG_ONLY_ACQ1=1 PDT_DEBUG_SKIP_SLOW=1 $B "$@" > gpurun_out/r02p_$name.json 2> gpurun_out/r02p_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02p_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["ms_per_step"],2))
except Exception as e: print("$name ERR", e, open("gpurun_out/r02p_$name.err").read()[-300:])
PY
}
run acq1_if1 --inflight 1
run acq1_if2 --inflight 2
run acq1_if4 --inflight 4
run acq1_if8 --inflight 8
run acq1_if16 --inflight 16 --groups 1
