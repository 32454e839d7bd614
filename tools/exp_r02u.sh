mkdir -p gpurun_out
B="python bench.py --steps 20 --no-e2e --no-cpu --no-single"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02u_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2), d["check"]["frames_decoded"])
except Exception as e: print("$name ERR", e)
PY
}
run base X=1
run skipslow PDT_DEBUG_SKIP_SLOW=1
run skip_gar_bits PDT_DEBUG_SKIP=k_gardner,k_bits
run skip_agc PDT_DEBUG_SKIP=k_agc_core
run skip_pll PDT_DEBUG_SKIP=k_pll_core,k_pll_fix_par
run only_acq PDT_DEBUG_SKIP=k_pll_core,k_pll_fix_par,k_front,k_agc_core,k_gardner,k_bits
run skipslow_only_acq PDT_DEBUG_SKIP_SLOW=1 PDT_DEBUG_SKIP=k_pll_core,k_pll_fix_par,k_front,k_agc_core,k_gardner,k_bits
EXTRA="--inflight 12" run if12 X=1
EXTRA="--inflight 6" run if6 X=1
python tools/timeline_inflight.py --inflight 8 > gpurun_out/r02u_tl8.txt 2>&1
