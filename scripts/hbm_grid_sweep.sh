#!/bin/bash
# HBM pass kernel: persistent-grid size (CTAs per SM) sweep on configs 5 and 4.
for m in 2 3 4 6; do
  for w in c5 c4; do
    OQRL_HBM_GRID_MULT=$m python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('grid_mult', $m, '$w', round(d['ms_per_step'], 4))"
  done
done
