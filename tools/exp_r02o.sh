mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-single --steps 20"
run() { name=$1; shift; env $ENVV $B "$@" > gpurun_out/r02o_$name.json 2> gpurun_out/r02o_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02o_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), [(k["kernel"],round(k["ms"],2)) for k in d["kernels"][:8]])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02o_$name.err").read()[-300:])
PY
}
ENVV="X=1" run cur_default
ENVV="PDT_LIB_VARIANT=_prev" run prev_default
ENVV="X=1" run cur_g3_if4 --groups 3 --inflight 4
ENVV="PDT_LIB_VARIANT=_prev" run prev_g3_if4 --groups 3 --inflight 4
ENVV="X=1" run cur_default_again
python tools/time_dropin.py --out gpurun_out/r02o_dropin.json
