#!/bin/bash
# HBM-resident class on the GPU box: parity tests, configs 5 and 4, and the phase timers of the pass kernel.
# usage: scripts/hbm_check.sh <tag>
tag=${1:-x}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hbm or c5 or c4" 2>&1 | tail -3
for w in c5 c4; do
  steps=50; [ $w = c4 ] && steps=3
  timeout 300 python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2_${w}_${tag}.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_${w}_${tag}.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
OQRL_HBM_DEBUG=2 timeout 200 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "hbm pass" | sed -n 20,24p
OQRL_HBM_DEBUG=2 timeout 200 python bench.py --workload c4 --population 64 --transitions 32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "hbm pass" | sed -n 5,6p
