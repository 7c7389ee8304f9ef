#!/bin/bash
# Scaling sweep on one 8-GPU box: config 2 weak scaling (4096 candidates per GPU) and config 3 strong scaling
# (16384 candidates in total, T = 256).  Usage: scripts/scale_sweep.sh <max_gpus> ; one JSON line per run -> gpurun_out/scale_*.jsonl
MAXN=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/scale_c2.jsonl; : > gpurun_out/scale_c3.jsonl
for N in 1 2 4 8; do
  [ $N -gt $MAXN ] && break
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N))"; fi
  $L bench.py --gpus $N --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/scale_c2.jsonl
  $L bench.py --gpus $N --workload c3 --population $((16384/N)) --transitions 256 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/scale_c3.jsonl
done
python - <<'PY'
import json
for f in ("gpurun_out/scale_c2.jsonl","gpurun_out/scale_c3.jsonl"):
    base=None
    for ln in open(f):
        try: d=json.loads(ln)
        except Exception: continue
        base=base or d["value"]
        print(f.split('/')[-1], d["n_gpus"], "%.3e evals/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], "eff %.3f"%(d["value"]/(base*d["n_gpus"]) if "c2" in f else d["value"]/(base*d["n_gpus"])), d["config"].get("collective"))
PY
