# multi-GPU evidence: default bench under torchrun at N ranks (N = $1)
N=$1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-single > gpurun_out/r02w_n${N}_bench.json 2> gpurun_out/r02w_n${N}_bench.err
if [ "$N" = "8" ]; then
  $T bench.py --gpus 8 --stream 1000000000 --fs 2000000 --steps 6 > gpurun_out/r02w_n8_stream_2M.json 2> gpurun_out/r02w_n8_stream_2M.err
  $T bench.py --gpus 8 --stream 1000000000 --fs 2000000 --steps 18 --inflight 6 > gpurun_out/r02w_n8_stream_2M_if6.json 2> gpurun_out/r02w_n8_stream_2M_if6.err
  $T bench.py --gpus 8 --stream 1000000000 --fs 250000 --steps 6 > gpurun_out/r02w_n8_stream_250k.json 2> gpurun_out/r02w_n8_stream_250k.err
fi
for f in gpurun_out/r02w_n${N}_*.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", round(d["value"]), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), (d.get("e2e_pcm16") or {}).get("value"), (d.get("strong_scaling") or {}).get("value"), d.get("per_rank",{}).get("ms_per_step"), d.get("check",{}).get("counter_breaks"))
except Exception as e: print("$f ERR", e)
PY
done
tail -c 200 gpurun_out/r02w_n${N}_bench.err
