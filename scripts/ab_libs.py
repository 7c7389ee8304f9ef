"""Kernel time of several builds of the library on one box (clocks differ between boxes and calls).

usage: python scripts/ab_libs.py P,T[,L] [P,T ...] -- lib.so [lib.so ...]     ("tree" = the in-tree library)
Each library runs in its own process (OQRL_B200_LIB); three interleaved rounds, median of 20 L2-flushed launches each.
"""
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
    rng = np.random.default_rng(0)
    n_rows = 4096
    ds = engine.DeviceDataset(rng.normal(size=(n_rows, 4)).astype(np.float32), rng.normal(size=(n_rows, 4)).astype(np.float32),
                              rng.integers(0, 2, n_rows).astype(np.int32), np.ones(n_rows, np.float32),
                              (rng.random(n_rows) < 0.05).astype(np.float32))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for case in json.loads(os.environ["OQRL_AB_WORKER"]):
        P, T = case[0], case[1]
        spec = CircuitSpec(4, case[2] if len(case) > 2 else 2, 2, 4, 0, 0, 0)
        params = torch.tensor(rng.normal(size=(P, spec.n_params)).astype(np.float32), device="cuda")
        idx = torch.tensor(rng.integers(0, n_rows, T).astype(np.int32), device="cuda")
        for _ in range(3):
            engine.population_fitness(spec, params, ds, idx)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.population_fitness(spec, params, ds, idx)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out[",".join(map(str, case))] = round(float(np.median(ts)), 2)
        # the host-buffer entry point (pinned parameters / indices in, pinned fitness out), wall clock around the call
        import time
        h_params = params.cpu().pin_memory()
        h_idx = idx.cpu().pin_memory()
        h_out = torch.empty(P, dtype=torch.float32).pin_memory()
        for _ in range(5):
            engine.population_fitness_host(spec, h_params, ds, h_idx, h_out)
        hs = []
        for _ in range(40):
            t0 = time.perf_counter()
            engine.population_fitness_host(spec, h_params, ds, h_idx, h_out)
            hs.append((time.perf_counter() - t0) * 1e6)
        out[",".join(map(str, case)) + ":host"] = round(float(np.median(hs)), 2)
    print(json.dumps(out))
    sys.exit(0)

args = sys.argv[1:]
cut = args.index("--")
cases = [[int(x) for x in a.split(",")] for a in args[:cut]]
libs = args[cut + 1:]
res = {lib: [] for lib in libs}
for _ in range(3):
    for lib in libs:
        env = dict(os.environ, OQRL_AB_WORKER=json.dumps(cases))
        if lib != "tree":
            env["OQRL_B200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(lib, "FAILED", r.stderr[-400:])
            continue
        res[lib].append(json.loads(r.stdout.strip().splitlines()[-1]))
for lib in libs:
    if res[lib]:
        print(os.path.basename(lib), {k: float(np.median([d[k] for d in res[lib]])) for k in res[lib][0]}, "us")
