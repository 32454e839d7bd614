mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02h_pytest.log; tail -2 gpurun_out/r02h_pytest.log
B="python bench.py --no-e2e --no-cpu --no-single --steps 16"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02h_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"], [(k["kernel"],round(k["ms"],1)) for k in d["kernels"][:6]])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02h_$name.err").read()[-300:])
PY
}
ONLY=k_pll_core,k_pll_fix_par,k_front,k_agc_core,k_gardner,k_bits
run rot_full X=1
run rot_onlyacq PDT_DEBUG_SKIP=$ONLY
run rot_skipslow PDT_DEBUG_SKIP_SLOW=1
run w4_full PDT_LIB_VARIANT=_w4
run w4_skipslow PDT_LIB_VARIANT=_w4 PDT_DEBUG_SKIP_SLOW=1
EXTRA="--inflight 6" run w4_full_if6 PDT_LIB_VARIANT=_w4
PDT_LIB_VARIANT=_w4 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
