mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02d_pytest.log; tail -3 gpurun_out/r02d_pytest.log
B="python bench.py --steps 12 --no-e2e --no-cpu --no-single"
run() { name=$1; shift; env "$@" $B > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],2))
except Exception as e: print("$name ERR", e)
PY
}
run base X=1
run lscap4 PDT_LS_CAP=4
run lscap1 PDT_LS_CAP=1
run acq96 PDT_ACQ0_THREADS=96
run acq128 PDT_ACQ0_THREADS=128
run groups1 PDT_MAX_GROUPS=1
run groups5 PDT_MAX_GROUPS=5
run acq96_ls4 PDT_ACQ0_THREADS=96 PDT_LS_CAP=4
python tools/timeline_inflight.py --inflight 1 > gpurun_out/r02d_tl1.txt 2>&1
python tools/timeline_inflight.py --inflight 4 > gpurun_out/r02d_tl4.txt 2>&1
