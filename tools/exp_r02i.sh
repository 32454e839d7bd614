mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-single --steps 16"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02i_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"], [(k["kernel"],round(k["ms"],1)) for k in d["kernels"][:4]])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02i_$name.err").read()[-300:])
PY
}
ONLY=k_pll_core,k_pll_fix_par,k_front,k_agc_core,k_gardner,k_bits
for v in _b128t128 _b128t224 _b64t128; do
  run full$v PDT_LIB_VARIANT=$v
  run onlyacq$v PDT_LIB_VARIANT=$v PDT_DEBUG_SKIP=$ONLY
done
EXTRA="--inflight 6" run full_b128t128_if6 PDT_LIB_VARIANT=_b128t128
EXTRA="--inflight 8" run full_b128t128_if8 PDT_LIB_VARIANT=_b128t128
EXTRA="--inflight 8" run full_b64t128_if8 PDT_LIB_VARIANT=_b64t128
PDT_LIB_VARIANT=_b128t128 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
PDT_LIB_VARIANT=_b64t128 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
