#!/bin/bash
# r02z7: tile-length sweep of the two time-tiled recurrences (PLL T/W, AGC T/(W/2)): redundant warm-up work against tile latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # pll_frac agc_halves
  PDT_AGC_TILE_HALVES=$2 timeout 300 python bench.py --no-cpu --no-single --no-e2e --steps 24 --pll-tile-frac $1 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = {r['kernel']: r['ms'] for r in b['kernels']}
print(json.dumps({'pll_tile_frac': $1, 'agc_tile_halves': $2, 'ms_per_step': round(b['ms_per_step'], 3), 'value': round(b['value']), 'k_pll_core': k.get('k_pll_core'), 'k_agc_core': k.get('k_agc_core'), 'frames': b['check']['frames_decoded'], 'spec': b['check']['speculation']}))"
}
IFS=","; for cfg in ${CFGS:-1 1,2 1,1 2,2 2,2 4,3 4}; do IFS=" "; run $cfg; IFS=","; done | tee -a gpurun_out/r02z7_tile_sweep.jsonl
