#!/bin/bash
for PT in "4096 1024" "4096 4096" "4096 16384" "2072 1024" "2072 8192" "4144 1024" "8288 1024" "16576 1024"; do
  set -- $PT
  python bench.py --workload c2 --population $1 --transitions $2 --no-cpu-baseline --no-also --steps 100 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('P', $1, 'T', $2, round(d['value'] / 1e9, 2), 'G evals/s', round(d['ms_per_step'], 4), 'ms', 'kern', d['roofline'].get('kernel_ms'), 'frac', round(d['roofline']['frac'], 3))"
done
