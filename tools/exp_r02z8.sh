#!/bin/bash
# r02z8: k_front1 as a persistent kernel (PDT_FRONT_PERSIST = resident CTAs per SM) against one CTA per tile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PDT_FRONT_PERSIST=7 timeout 900 python -m pytest tests -m gpu -q -x -k "poes or tiled or batch or config" 2>&1 | tail -3
run() {
  PDT_FRONT_PERSIST=$1 timeout 300 python bench.py --no-cpu --no-single --no-e2e --steps 16 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = {r['kernel']: (r['ms'], r.get('ms_per_launch')) for r in b['kernels']}
print(json.dumps({'front_persist': $1, 'ms_per_step': round(b['ms_per_step'], 3), 'k_front': k.get('k_front'), 'roofline_frac': round(b['roofline']['frac'], 4), 'frames': b['check']['frames_decoded']}))"
}
for p in 0 7 5 0 7; do run $p; done | tee gpurun_out/r02z8_front_persist.jsonl
