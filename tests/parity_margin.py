"""Worst relative error of the reference-agent kernels against the fp64 oracle over many random candidates
(test infrastructure, run by hand on the GPU box: python tests/parity_margin.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oqrl_b200 import CircuitSpec, engine
from oracle import c_oracle, pqc_oracle as po
s = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cartpole_v5_sample.npz"))
rng = np.random.default_rng(0)
for L in (2, 3, 4):
    for ent in (0, 1):
        spec = po.Spec(4, L, 2, 4, ent)
        ps = CircuitSpec(4, L, 2, 4, ent)
        params = (rng.random((4096, spec.n_params)) * 6.3 - 3.1).astype(np.float32)      # full angle range
        x = s["sample_obs"][rng.integers(0, 4096, 64)]
        q = engine.qvalues(ps, params, x).cpu().numpy()
        ref = c_oracle.qvalues(spec, params, x)
        err = np.abs(q - ref) / np.maximum(np.abs(ref), 1.0)
        print("L", L, "ent", ent, "max rel err", float(err.max()), "mean", float(err.mean()))
