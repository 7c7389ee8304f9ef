#!/bin/bash
# Config-2 kernel across population sizes (whole / split work-item mix).
for P in 2500 3000 4096 5000 8192; do
  python bench.py --workload c2 --population $P --no-cpu-baseline --steps 300 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('P', $P, round(d['value'] / 1e9, 2), 'G evals/s', round(d['ms_per_step'], 4), 'ms', 'frac', round(d['roofline']['frac'], 3))"
done
