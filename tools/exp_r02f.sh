mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02f_pytest.log; tail -2 gpurun_out/r02f_pytest.log
B="python bench.py --no-e2e --no-cpu --no-single"
run() { name=$1; shift; $B "$@" > gpurun_out/r02f_$name.json 2> gpurun_out/r02f_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02f_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02f_$name.err").read()[-300:])
PY
}
run if4_k20 --steps 20 --inflight 4
run if6_k20 --steps 20 --inflight 6
run if8_k20 --steps 20 --inflight 8
run if10_k20 --steps 20 --inflight 10
run if8_k60 --steps 60 --inflight 8
run if12_k60 --steps 60 --inflight 12
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
