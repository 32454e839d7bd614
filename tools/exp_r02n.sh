mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --no-e2e --no-cpu --no-single --steps 20 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1]); print("bench", round(d["value"]), round(d["ms_per_step"],2), d["check"], [(k["kernel"],round(k["ms"],2)) for k in d["kernels"][:9]])
PY
python bench.py --stream 1000000000 --fs 2000000 --steps 6 > gpurun_out/r02n_stream_2M.json 2> gpurun_out/r02n_stream_2M.err; cut -c1-200 gpurun_out/r02n_stream_2M.json; python -c "
import json; d=json.loads(open('gpurun_out/r02n_stream_2M.json').read().strip().splitlines()[-1]); print(d['check'])"
python bench.py --stream 1000000000 --fs 2000000 --steps 6 --pll-tile-frac 0.25 > gpurun_out/r02n_stream_2M_t025.json 2> gpurun_out/r02n_stream_2M_t025.err; cut -c1-200 gpurun_out/r02n_stream_2M_t025.json
python bench.py --stream 1000000000 --fs 2000000 --steps 12 --inflight 6 > gpurun_out/r02n_stream_2M_if6.json 2> gpurun_out/r02n_stream_2M_if6.err; cut -c1-200 gpurun_out/r02n_stream_2M_if6.json
python tools/timeline_stream.py --fs 2000000 > gpurun_out/r02n_tl_stream_2M.txt 2>&1; tail -40 gpurun_out/r02n_tl_stream_2M.txt
