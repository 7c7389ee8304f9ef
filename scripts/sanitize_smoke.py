"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck), one tool per run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oqrl_b200 import CircuitSpec, engine  # noqa: E402

s = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_v5_sample.npz"))
ds = engine.DeviceDataset(s["sample_obs"][:512], s["sample_next_obs"][:512], s["sample_act"][:512], s["sample_rew"][:512], s["sample_term"][:512])
rng = np.random.default_rng(0)
idx = rng.choice(512, 70, replace=False).astype(np.int32)
for spec, P in ((CircuitSpec(), 37), (CircuitSpec(4, 3, 2, 4, 1, 1, 0), 5), (CircuitSpec(3, 2, 2, 3), 5), (CircuitSpec(6, 2, 2, 4, 0, 1, 1), 3),
                (CircuitSpec(8, 3, 2, 4, 0, 1, 1), 3), (CircuitSpec(10, 2, 2, 4, 1, 0, 1), 2), (CircuitSpec(12, 2, 2, 4, 0, 1, 1), 2),
                (CircuitSpec(8, 2, 2, 4, 0, 1, 1, general_kernel=1), 2), (CircuitSpec(4, 3, 2, 4), 9), (CircuitSpec(14, 2, 2, 4, 0, 0, 1), 1)):
    p = rng.normal(size=(P, spec.n_params)).astype(np.float32)
    x = s["sample_obs"][:5, :spec.n_feat]
    if spec.n_feat == 4:
        f = engine.population_fitness(spec, p, ds, idx[: 70 if spec.n_qubits < 14 else 3])
        assert torch.isfinite(f).all()
    q = engine.qvalues(spec, p, x)
    assert torch.isfinite(q).all()
buf = torch.zeros(64, device="cuda")
engine.population_fitness_scatter(CircuitSpec(), rng.normal(size=(20, 24)).astype(np.float32), ds, idx, [buf.data_ptr()], 11)
pop = torch.tensor(rng.normal(size=(25, 24)).astype(np.float32), device="cuda")
engine.ga_breed(pop, np.arange(5, dtype=np.int32), 2, rng.integers(0, 5, (12, 2)).astype(np.int32), rng.integers(0, 2, (12, 24)).astype(np.uint8), rng.normal(size=(12, 2, 24)))
fit = torch.tensor(rng.normal(size=25).astype(np.float32), device="cuda")
order = engine.ga_rank(fit, 5)
engine.ga_breed_random(pop, order, 2, 5, 0.7, 0.1, seed=1, generation=0)
flags = torch.zeros(2, dtype=torch.int32, device="cuda")
engine.peer_barrier([flags.data_ptr()], 0, 1)
torch.cuda.synchronize()
print("sanitize smoke ok")
