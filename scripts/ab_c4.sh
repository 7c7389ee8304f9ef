#!/bin/bash
# Config-4-shaped proxy (16 qubits, 2 layers, P = 128 x T = 128) for several builds of the library on one box.
for r in 1 2; do
  for lib in "" "$@"; do
    OQRL_B200_LIB=$lib python bench.py --workload c4 --population 128 --transitions 128 --no-cpu-baseline --no-also --steps 8 --warmup 2 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${lib:-tree}', round(d['value']), 'evals/s', round(d['ms_per_step'], 3), 'ms')"
  done
done
