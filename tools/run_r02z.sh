#!/bin/bash
# r02z: exact engine / legacy PLL through the block runner (pdt_pll_pipe.cuh) - parity, then ARGOS, live mode, drop-in, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02z_pytest.txt
cat gpurun_out/r02z_pytest.txt
python bench.py --mode argos --steps 6 --warmup 3 > gpurun_out/r02z_argos.json 2> gpurun_out/r02z_argos.err; tail -c 600 gpurun_out/r02z_argos.json
python tools/live_latency.py gpurun_out/r02z_live_latency.json
python tools/time_dropin.py --out gpurun_out/r02z_dropin.json
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; python - <<'PY'
import json
b = json.loads(open("gpurun_out/r02z_bench.json").read().strip().splitlines()[-1])
print("value", b["value"], "ms", b["ms_per_step"], "e2e", b.get("e2e", {}).get("value"), "roofline", b["roofline"]["frac"], b["roofline"]["kernel_ms"])
for r in b["kernels"]:
    print(r["kernel"], r["ms"], r.get("ms_per_launch"))
PY
