"""Kernel time of the shared-memory class (config-3-shaped circuits: re-uploading, tiled features) for several builds of
the library on one box:  python scripts/ab_c3.py [P T] -- lib.so ...   ("tree" = the in-tree library)."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if os.environ.get("OQRL_AB_WORKER"):
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from oqrl_b200 import CircuitSpec, engine
    P, T = json.loads(os.environ["OQRL_AB_WORKER"])
    rng = np.random.default_rng(0)
    n_rows = 4096
    ds = engine.DeviceDataset(rng.normal(size=(n_rows, 4)).astype(np.float32), rng.normal(size=(n_rows, 4)).astype(np.float32),
                              rng.integers(0, 2, n_rows).astype(np.int32), np.ones(n_rows, np.float32),
                              (rng.random(n_rows) < 0.05).astype(np.float32))
    out = {}
    for n, L in ((8, 8), (6, 4), (10, 4), (12, 2)):
        spec = CircuitSpec(n, L, 2, 4, 0, 1, 1)
        p = max(148, P * 256 // (1 << n))
        params = torch.tensor(rng.normal(size=(p, spec.n_params)).astype(np.float32), device="cuda")
        idx = torch.tensor(rng.integers(0, n_rows, T).astype(np.int32), device="cuda")
        for _ in range(2):
            engine.population_fitness(spec, params, ds, idx)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.population_fitness(spec, params, ds, idx)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[f"n{n}L{L}P{p}"] = round(float(np.median(ts)), 3)
    print(json.dumps(out))
    sys.exit(0)

args = sys.argv[1:]
cut = args.index("--")
P, T = (int(args[0]), int(args[1])) if cut >= 2 else (2048, 256)
libs = args[cut + 1:]
for _ in range(2):
    for lib in libs:
        env = dict(os.environ, OQRL_AB_WORKER=json.dumps([P, T]))
        if lib != "tree":
            env["OQRL_B200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.returncode == 0 else "FAILED " + r.stderr[-300:], "ms")
