#!/bin/bash
# Config-2 kernel time of two builds of the library in one box (clocks differ between boxes): the tree's library and
# the one named by $1 (e.g. ab_libs/liboqrl_b200_r2base.so, built from an older commit).  Interleaved, three rounds.
for r in 1 2 3; do
  for lib in "" "$1"; do
    OQRL_B200_LIB=$lib python bench.py --workload c2 --no-cpu-baseline --no-also --steps 300 ${@:2} 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${lib:-tree}', round(d['value'] / 1e9, 2), 'G evals/s', round(d['ms_per_step'], 4), 'ms', 'kern', round(d['roofline'].get('kernel_ms', 0), 5), 'frac', round(d['roofline']['frac'], 3), 'sm_mhz', d.get('clocks', {}).get('sm_mhz'))"
  done
done
