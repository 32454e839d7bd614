#!/bin/bash
# r02z3: slow-capture launch of k_front1 compacted; full parity, default bench, ncu --set full of the exact engine's kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02z3_pytest.txt; cat gpurun_out/r02z3_pytest.txt
timeout 900 python bench.py > gpurun_out/r02z3_bench.json 2> gpurun_out/r02z3_bench.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/r02z3_bench.json").read().strip().splitlines()[-1])
print("value", b["value"], "ms", b["ms_per_step"], "e2e", b.get("e2e", {}).get("value"), "roofline", b["roofline"]["frac"], b["roofline"]["kernel_ms"], "check", b["check"]["frames_decoded"], b.get("cpu_baseline", {}).get("frame_bytes_check"))
for r in b["kernels"]:
    print(r["kernel"], r["ms"], r.get("ms_per_launch"))
PY
timeout 900 ncu --set full --clock-control none -k regex:k_chain_exact -s 1 -c 1 --csv --page raw --log-file gpurun_out/r02z3_ncu_chain_exact_argos_raw.csv python tools/chain_prof.py > /tmp/ncu_chain.log 2>&1; tail -2 /tmp/ncu_chain.log | cut -c1-300
