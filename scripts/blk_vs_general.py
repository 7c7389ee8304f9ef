"""A/B of the two shared-memory-class implementations (register-block pqc_blk.cu vs general pqc_smem.cu).

Prints kernel time per (n, L) for both; used to decide the dispatch in api.cu::use_blk.  GPU box only.
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


def time_ms(spec, P, T, reps=5):
    params = torch.tensor(rng.normal(size=(P, spec.n_params)).astype(np.float32), device="cuda")
    idx = torch.tensor(rng.integers(0, N_ROWS, T).astype(np.int32), device="cuda")
    for _ in range(2):
        engine.population_fitness(spec, params, ds, idx)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.population_fitness(spec, params, ds, idx)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


layers = [int(a) for a in sys.argv[1:]] or [6]
for L in layers:
    for n in range(5, 14):
        evals = 2 ** 27 // (2 ** n * max(L, 1))          # roughly constant work per case
        T = 256
        P = max(148, evals // T)
        spec = CircuitSpec(n, L, 2, 4, 0, 1, 1)
        name = engine.kernel_name(spec)
        a = time_ms(spec, P, T)
        b = time_ms(dataclasses.replace(spec, general_kernel=1), P, T)
        fl, _ = engine.work_model(spec)
        print(json.dumps({"n": n, "L": L, "P": P, "T": T, "kernel": name, "blk_ms": round(a, 3), "general_ms": round(b, 3),
                          "speedup": round(b / a, 3), "blk_tflops_executed": round(fl * P * T * 2 / a / 1e9, 1)}), flush=True)
