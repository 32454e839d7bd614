mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-single --steps 20"
run() { name=$1; shift; $B "$@" > gpurun_out/r02j_$name.json 2> gpurun_out/r02j_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02j_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02j_$name.err").read()[-300:])
PY
}
run g1_if4 --groups 1 --inflight 4
run g1_if6 --groups 1 --inflight 6
run g1_if8 --groups 1 --inflight 8
run g1_if10 --groups 1 --inflight 10
run g1_if12 --groups 1 --inflight 12
run g2_if6 --groups 2 --inflight 6
run g2_if8 --groups 2 --inflight 8
run g3_if6 --groups 3 --inflight 6
run g1_if12_k40 --groups 1 --inflight 12 --steps 40
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
