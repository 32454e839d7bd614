mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --no-single --steps 16"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02g_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"])
except Exception as e: print("$name ERR", e, open("gpurun_out/r02g_$name.err").read()[-300:])
PY
}
ONLY=k_pll_core,k_pll_fix_par,k_front,k_agc_core,k_gardner,k_bits
EXTRA="--inflight 8" run onlyacq_if8 PDT_DEBUG_SKIP=$ONLY
EXTRA="--inflight 4" run acq1only_if4 PDT_DEBUG_SKIP_SLOW=2
EXTRA="--inflight 8" run acq1only_if8 PDT_DEBUG_SKIP_SLOW=2
EXTRA="--inflight 4" run slowpipe_only_if4 PDT_DEBUG_SKIP_SLOW=3
EXTRA="--inflight 8" run slowpipe_only_if8 PDT_DEBUG_SKIP_SLOW=3
EXTRA="--inflight 8" run skipslow_if8 PDT_DEBUG_SKIP_SLOW=1
EXTRA="--inflight 8" run skipslow_if8_nogar PDT_DEBUG_SKIP_SLOW=1 PDT_DEBUG_SKIP=k_gardner,k_bits
EXTRA="--inflight 8" run skipslow_if8_noagc PDT_DEBUG_SKIP_SLOW=1 PDT_DEBUG_SKIP=k_agc_core,k_gardner,k_bits
EXTRA="--inflight 8" run skipslow_if8_nopll PDT_DEBUG_SKIP_SLOW=1 PDT_DEBUG_SKIP=k_pll_core,k_pll_fix_par
