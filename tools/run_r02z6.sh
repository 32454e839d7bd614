#!/bin/bash
# r02z6 (was r02z2): exact engine after the pipelined PLL runner + window staging: parity, accounting, ARGOS bench, live mode, drop-in
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02z6_pytest.txt; cat gpurun_out/r02z6_pytest.txt
timeout 200 python tools/chain_prof.py gpurun_out/r02z6_chain_prof.json
timeout 600 python bench.py --mode argos --steps 6 --warmup 3 > gpurun_out/r02z6_argos.json 2> gpurun_out/r02z6_argos.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/r02z6_argos.json").read().strip().splitlines()[-1])
print("ARGOS value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], b.get("check"), "cpu", b.get("cpu_baseline", {}).get("value"))
PY
timeout 300 python tools/live_latency.py gpurun_out/r02z6_live_latency.json
timeout 300 python tools/time_dropin.py --out gpurun_out/r02z6_dropin.json
