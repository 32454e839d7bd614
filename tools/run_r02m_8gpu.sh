# 8-GPU evidence run (one box): default bench (weak + strong leg), the same without the gather, configs[4] streams
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-single > gpurun_out/r02m_n8_bench.json 2> gpurun_out/r02m_n8_bench.err
$T bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-single --no-e2e --no-gather > gpurun_out/r02m_n8_bench_nogather.json 2> gpurun_out/r02m_n8_bench_nogather.err
$T bench.py --gpus 8 --stream 1000000000 --fs 2000000 --steps 6 > gpurun_out/r02m_n8_stream_2M.json 2> gpurun_out/r02m_n8_stream_2M.err
$T bench.py --gpus 8 --stream 1000000000 --fs 2000000 --steps 6 --pll-tile-frac 0.25 > gpurun_out/r02m_n8_stream_2M_t025.json 2> gpurun_out/r02m_n8_stream_2M_t025.err
$T bench.py --gpus 8 --stream 1000000000 --fs 250000 --steps 6 > gpurun_out/r02m_n8_stream_250k.json 2> gpurun_out/r02m_n8_stream_250k.err
for f in gpurun_out/r02m_n8_*.json; do echo $f; cut -c1-400 $f; done
tail -c 300 gpurun_out/r02m_n8_*.err
