#!/bin/bash
# Configs 5 and 4 with the tree's library and the one named by $1, interleaved on one box.
for r in 1 2; do
  for lib in "" "$1"; do
    for w in c5 c4; do
    OQRL_B200_LIB=$lib python bench.py --workload $w --no-cpu-baseline --no-also --steps 20 --warmup 3 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${lib:-tree}', '$w', round(d['value'], 1), 'evals/s', round(d['ms_per_step'], 4), 'ms', 'frac', round(d['roofline']['frac'], 3))"
    done
  done
done
