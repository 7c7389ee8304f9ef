"""Times the reference's OWN hot path -- GAOptimizer._evaluate_fitness (optim.py:64-173) over VQC.forward
(agent.py:48-61), the files of oracle/_ref/ copied unmodified from the reference -- behind the pennylane import shim
(tests/golden/shims/pennylane.py -> the float64 oracle; PennyLane itself is not installable offline).
BASELINE INFRASTRUCTURE: used by bench.py's CPU legs only, never by the product path.

What is and is not the reference here: the Python control flow, the per-sample QNode loop, the torch re-stacking and
the TD/MSE arithmetic are the reference's code; the circuit simulation inside each QNode call is the oracle's numpy
restatement, which is much cheaper than PennyLane's tape machinery -- so this rate is an UPPER bound of the real
reference's rate (kind "reference+shim").
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
SHIMS = os.path.join(ROOT, "tests", "golden", "shims")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF, f)) for f in ("optim.py", "agent.py", "utils.py", "config.py"))


def _import_reference():
    for p in (REF, SHIMS, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import agent as ref_agent          # noqa: E402  (the reference's files)
    import optim as ref_optim          # noqa: E402
    import utils as ref_utils          # noqa: E402
    return ref_agent, ref_optim, ref_utils


def fitness_rate(n_qubits, n_layers, n_out, params, obs, act, rew, next_obs, term, seconds, threads=1):
    """circuit-evals/s of the reference loop: candidates of `params` are evaluated one after another on the whole
    batch until `seconds` have passed.  Returns (rate, candidates evaluated, fitness values)."""
    import numpy as np
    import torch
    torch.set_num_threads(threads)     # the reference is single-threaded by construction (agent.py:51, optim.py:204)
    ref_agent, ref_optim, ref_utils = _import_reference()
    assert str(ref_utils.device) == "cpu", "run the reference arm with CUDA hidden (CUDA_VISIBLE_DEVICES='')"
    rng = np.random.default_rng(0)
    model = ref_agent.VQC(n_qubits, n_out, n_layers=n_layers, rng=rng)
    ga = ref_optim.GAOptimizer(model, population_size=1, num_generations=1, rng=rng)
    batch = [ref_utils.Experience(torch.tensor(obs[i], dtype=torch.float32), torch.tensor(int(act[i]), dtype=torch.int64),
                                  torch.tensor(float(rew[i]), dtype=torch.float32), torch.tensor(next_obs[i], dtype=torch.float32),
                                  torch.tensor(float(term[i]), dtype=torch.float32)) for i in range(len(obs))]
    fits, t0, p = [], time.perf_counter(), 0
    while True:
        fits.append(ga._evaluate_fitness([torch.tensor(params[p % len(params)])], None, batch))
        p += 1
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return 2.0 * p * len(batch) / el, p, fits


def fitness_rate_subprocess(n_qubits, n_layers, n_out, params, obs, act, rew, next_obs, term, seconds):
    """fitness_rate() in a child process whose CUDA devices are hidden: the reference's utils.py picks `cuda` whenever
    torch sees one (utils.py:9), and its files must run on the CPU exactly as written."""
    import json
    import subprocess
    import tempfile
    import numpy as np
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.npz")
        np.savez(path, params=params, obs=obs, act=act, rew=rew, next_obs=next_obs, term=term)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        out = subprocess.run([sys.executable, os.path.abspath(__file__), path, str(n_qubits), str(n_layers), str(n_out), str(seconds)],
                             capture_output=True, text=True, env=env, timeout=seconds * 10 + 120)
    if out.returncode != 0:
        raise RuntimeError("reference files failed to run: " + out.stderr[-800:])
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    return rec["rate"], rec["candidates"], rec["fitness"]


if __name__ == "__main__":
    import json
    import numpy as np
    d = np.load(sys.argv[1])
    n, L, no, secs = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
    rate, cands, fits = fitness_rate(n, L, no, d["params"], d["obs"], d["act"], d["rew"], d["next_obs"], d["term"], secs)
    print(json.dumps({"rate": rate, "candidates": cands, "fitness": fits[:8]}))
