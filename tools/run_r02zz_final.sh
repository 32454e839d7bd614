# final single-GPU evidence run of round 2 (after the exact-engine work and the compact slow-capture launch of k_front1)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02zz_pytest.log; cat gpurun_out/r02zz_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zz_bench_reference_arm.json 2>/dev/null
python bench.py > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err; tail -c 300 gpurun_out/r02zz_bench.err
python bench.py --fs 50000 --captures 512 --no-cpu --no-single > gpurun_out/r02zz_bench_L3.json 2> gpurun_out/r02zz_bench_L3.err
python bench.py --fs 18750 --captures 256 --steps 8 --inflight 3 --no-cpu --no-single > gpurun_out/r02zz_bench_L8.json 2> gpurun_out/r02zz_bench_L8.err
python bench.py --stream 1000000000 --fs 250000 --steps 6 > gpurun_out/r02zz_stream_250k_1G.json 2> gpurun_out/r02zz_stream_250k_1G.err
python bench.py --mode argos --steps 10 > gpurun_out/r02zz_argos.json 2> gpurun_out/r02zz_argos.err
python bench.py --mode argos --impl reference --steps 3 --warmup 1 > gpurun_out/r02zz_argos_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02zz_launches.csv python bench.py --steps 4 --warmup 3 --inflight 4 --no-e2e --no-cpu --no-single > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_front1" -s 10 -c 4 -o /tmp/r02zz_full python bench.py --steps 1 --warmup 3 --inflight 1 --no-e2e --no-cpu --no-single > gpurun_out/r02zz_ncu_full.log 2>&1
ncu -i /tmp/r02zz_full.ncu-rep --page raw --csv > gpurun_out/r02zz_ncu_front1_raw.csv 2>/dev/null
for f in bench bench_reference_arm bench_L3 bench_L8 stream_250k_1G argos argos_reference_arm; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02zz_$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"]), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"))
except Exception as e: print("$f ERR", e)
PY
done
