"""End-to-end GAOptimizer.optimize() timing (scope row f.1): list-based mirror vs device-resident population.

Includes everything a generation costs on each path: fitness kernel, selection, the reference's
RNG draws, and breeding.  Prints one JSON line per (optimizer, population)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oqrl_b200 import DeviceGAOptimizer, GAOptimizer, VQC, engine  # noqa: E402
from oqrl_b200.dataset import IndexBatch  # noqa: E402

sample = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_v5_sample.npz"))
ds = engine.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
for cls, kw in ((GAOptimizer, {}), (DeviceGAOptimizer, {}), (DeviceGAOptimizer, {"exact_rng_stream": False})):
    for P, T, G in ((25, 16, 10), (4096, 1024, 3)):
        if cls is GAOptimizer and P > 1024:
            G = 1
        rng = np.random.default_rng(0)
        model = VQC(4, 2, n_layers=2, rng=rng)
        ga = cls(model, population_size=P, num_generations=G, rng=rng, **kw)
        batch = IndexBatch(ds, np.random.default_rng(1).choice(4096, T, replace=False))
        ga.optimize(None, batch)            # warm-up (allocations, first launches)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3 if P <= 1024 or cls is DeviceGAOptimizer else 1
        for _ in range(reps):
            ga.optimize(None, batch)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"optimizer": cls.__name__ + ("" if kw.get("exact_rng_stream", True) else " (in-kernel RNG, no stream parity)"), "population": P, "transitions": T, "generations": G,
                          "s_per_optimize": dt, "ms_per_generation": 1e3 * dt / G,
                          "circuit_evals_per_s_end_to_end": 2.0 * P * T * G / dt}), flush=True)
