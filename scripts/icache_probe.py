"""Instruction-fetch probe: packed-FMA throughput versus unrolled loop-body size and resident warps.

Decides how large the unrolled hot loops of the circuit kernels may be (profiles/r1_summary.md).
Run on the GPU box:  python scripts/icache_probe.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oqrl_b200 import engine  # noqa: E402

BODY = [256, 384, 512, 768, 1024, 1536, 2048, 3072]
out = {}
for warps in (1, 2, 3, 4, 8):
    row = {}
    for v, body in enumerate(BODY):
        iters = max(64, (1 << 21) // body)
        row[f"{body * 16 // 1024}KB"] = round(engine.fma_peak_tflops(32 + v + (warps << 8), iters), 1)
    out[f"{warps} warps/SMSP"] = row
    print(warps, row, flush=True)
print(json.dumps(out))
