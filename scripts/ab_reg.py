"""A/B of the register-class fitness kernel: batch staged in shared memory vs direct loads (spec.direct_loads).

Interleaved repetitions in one process (clocks differ between boxes and calls); prints the median kernel time per
(P, T) for both.  GPU box only.
"""
import dataclasses
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oqrl_b200 import CircuitSpec, engine  # noqa: E402

rng = np.random.default_rng(0)
N_ROWS = 4096
obs = rng.normal(size=(N_ROWS, 4)).astype(np.float32)
nxt = rng.normal(size=(N_ROWS, 4)).astype(np.float32)
act = rng.integers(0, 2, N_ROWS).astype(np.int32)
rew = np.ones(N_ROWS, np.float32)
term = (rng.random(N_ROWS) < 0.05).astype(np.float32)
ds = engine.DeviceDataset(obs, nxt, act, rew, term)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def time_ms(specs, P, T, reps=15):
    params = torch.tensor(rng.normal(size=(P, specs[0].n_params)).astype(np.float32), device="cuda")
    idx = torch.tensor(rng.integers(0, N_ROWS, T).astype(np.int32), device="cuda")
    for s in specs:
        for _ in range(3):
            engine.population_fitness(s, params, ds, idx)
    torch.cuda.synchronize()
    ts = [[] for _ in specs]
    for _ in range(reps):
        for i, s in enumerate(specs):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.population_fitness(s, params, ds, idx)
            e1.record()
            torch.cuda.synchronize()
            ts[i].append(e0.elapsed_time(e1))
    return [float(np.median(t)) for t in ts]


spec = CircuitSpec(4, int(os.environ.get("L", "2")), 2, 4, 0, 0, 0)
cases = [(4096, 1024), (2368, 1024), (4736, 1024), (16576, 1024), (4096, 2048), (1024, 1024), (256, 256), (4096, 64)]
for P, T in cases:
    a, b = time_ms([spec, dataclasses.replace(spec, direct_loads=1)], P, T)
    print(json.dumps({"P": P, "T": T, "staged_ms": round(a, 4), "direct_ms": round(b, 4), "speedup": round(b / a, 3),
                      "staged_Gevals": round(2 * P * T / a / 1e6, 1)}), flush=True)
