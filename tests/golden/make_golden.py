"""Regenerates the fixtures in tests/golden/ (run in the BUILD container only).

Needs /root/reference (read-only mount).  Nothing under tests/ reads the
reference at test time; the GPU box only sees the committed outputs:

  cartpole_v5_sample.npz   rows of offline_cartpole_test_v5.hdf5 read with
                           oqrl_b200.hdf5_subset (head 64 rows + a seeded
                           4096-row sample) and the Appendix-C checksums
  kat_appendix_b.json      SURVEY.md Appendix B known-answer vectors
  ga_trajectory_ref.npz    produced by importing the reference's OWN
                           /root/reference/src/optim.py + utils.py UNMODIFIED
                           and driving GAOptimizer.optimize with a model whose
                           forward is the float64 oracle (PennyLane is not
                           installable, so agent.py itself cannot be imported)
  golden_fitness.npz       oracle outputs for a grid of specs (regression pins
                           for the CUDA kernels)
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oqrl_b200 import hdf5_subset  # noqa: E402
from oracle import pqc_oracle as po  # noqa: E402


def load_ref_dataset(name="offline_cartpole_test_v5.hdf5"):
    path = os.path.join(REF, name)
    f = hdf5_subset.File(path)
    cols = {k: np.array(f[k]) for k in ("observations", "actions", "rewards", "next_observations", "terminals")}
    sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
    return cols, sha


def make_sample():
    cols, sha = load_ref_dataset()
    n = cols["observations"].shape[0]
    idx = np.random.default_rng(0).choice(n, size=4096, replace=False)   # dataset.py:58 with seed 0
    out = {
        "sha256": np.array(sha),
        "n_rows": np.array(n),
        "sum_actions": np.array(int(cols["actions"].sum())),
        "sum_terminals": np.array(int(cols["terminals"].sum())),
        "sum_obs_cols": cols["observations"].sum(axis=0, dtype=np.float64),
        "sample_idx": idx.astype(np.int64),
    }
    for tag, sel in (("head", np.arange(64)), ("sample", idx)):
        out[f"{tag}_obs"] = cols["observations"][sel]
        out[f"{tag}_next_obs"] = cols["next_observations"][sel]
        out[f"{tag}_act"] = cols["actions"][sel].reshape(-1).astype(np.int64)
        out[f"{tag}_rew"] = cols["rewards"][sel].astype(np.float64)
        out[f"{tag}_term"] = cols["terminals"][sel]
    np.savez_compressed(os.path.join(HERE, "cartpole_v5_sample.npz"), **out)
    return cols


def make_kat():
    pi = float(np.pi)
    kat = {
        "provenance": "SURVEY.md Appendix B (survey-derived; two independent float64 evaluators; NOT PennyLane output)",
        "basis_cases_n4_L2_w0": [
            {"x": [0, 0, 0, 0], "z": [1, 1, 1, 1]},
            {"x": [pi, 0, 0, 0], "z": [-1, -1, -1, 1]},
            {"x": [0, pi, 0, 0], "z": [-1, -1, 1, 1]},
            {"x": [0, 0, pi, 0], "z": [-1, -1, 1, -1]},
            {"x": [0, 0, 0, pi], "z": [1, -1, -1, -1]},
        ],
        "linspace_case": {"x": [0.1, -0.2, 0.3, -0.4], "w": "np.linspace(-1.2, 1.1, 24).astype(float32)",
                          "q": [0.1310773021, 0.1543603522]},
        "w0": "np.random.default_rng(0).random(24).astype(float32)",
        "w0_rows": {"0": [0.2848576149, 0.1965694125], "1": [0.3610721505, 0.1841487482],
                    "2": [0.2831204485, 0.1962711836], "3": [0.2037061973, 0.1664040913]},
        "w0_fitness_rows_0_15_gamma_0.99": -1.066803366503536,
        "sel_ranges": {"2,3": [1, 1, 1], "3,4": [1, 2, 1, 2], "5,5": [1, 2, 3, 4, 1], "8,8": [1, 2, 3, 4, 5, 6, 7, 1]},
    }
    json.dump(kat, open(os.path.join(HERE, "kat_appendix_b.json"), "w"), indent=1)


class OracleVQC(torch.nn.Module):
    """Stand-in for agent.py's VQC: same params / forward contract, oracle arithmetic."""

    def __init__(self, input_dim, output_dim, n_layers, rng):
        super().__init__()
        self.spec = po.Spec(n_qubits=input_dim, n_layers=n_layers, n_out=output_dim, n_feat=input_dim)
        init = rng.random(n_layers * input_dim * 3).astype(np.float32)      # agent.py:30
        self.params = torch.nn.Parameter(torch.tensor(init), requires_grad=False)

    def forward(self, x):
        q = po.qvalues_tensor(self.spec, self.params.detach().numpy(), x.detach().numpy())
        return torch.tensor(q)   # float64, like the QNode output


def make_ga_trajectory(cols):
    sys.path.insert(0, os.path.join(REF, "src"))
    import optim as ref_optim          # the reference's own file, unmodified
    import utils as ref_utils
    out = {}
    for seed in (0, 1):
        rng = np.random.default_rng(seed)                                   # utils.py:22 (without global seeding)
        policy = OracleVQC(4, 2, 2, rng)
        _target = OracleVQC(4, 2, 2, rng)                                    # agent.py:90 consumes the same stream
        _first = ref_optim.GAOptimizer(model=policy, rng=rng)                # agent.py:99
        ga = ref_optim.GAOptimizer(policy, rng=rng)                          # train.py:37
        # 16 transitions drawn from the committed 4096-row sample, so that the tests can rebuild the batch
        # without the reference data file
        sample_idx = np.random.default_rng(0).choice(100000, size=4096, replace=False)
        rows = sample_idx[rng.choice(4096, 16, replace=False)]
        batch = [ref_utils.Experience(torch.tensor(cols["observations"][r]),
                                      torch.tensor(cols["actions"][r].reshape(()), dtype=torch.int64),
                                      torch.tensor(cols["rewards"][r], dtype=torch.float32),
                                      torch.tensor(cols["next_observations"][r]),
                                      torch.tensor(float(cols["terminals"][r]), dtype=torch.float32)) for r in rows]
        init_pop = np.stack([ind[0].numpy().copy() for ind in ga.population])
        log = []
        orig = ga._evaluate_fitness

        def spy(ind, loss_fn, b, _orig=orig, _log=log):
            v = _orig(ind, loss_fn, b)
            _log.append(v)
            return v

        ga._evaluate_fitness = spy
        ga.optimize(None, batch)
        tag = f"seed{seed}_"
        out[tag + "rows"] = rows
        out[tag + "init_population"] = init_pop
        out[tag + "fitness_log"] = np.array(log).reshape(ga.num_generations, ga.population_size)
        out[tag + "final_population"] = np.stack([ind[0].numpy() for ind in ga.population])
        out[tag + "final_model_params"] = policy.params.detach().numpy().copy()
        out[tag + "best_individual"] = ga.best_individual[0].numpy().copy()
        out[tag + "interaction_count"] = np.array(ga.interaction_count)
        out[tag + "rng_probe"] = np.array(rng.random())
    np.savez_compressed(os.path.join(HERE, "ga_trajectory_ref.npz"), **out)


GOLDEN_SPECS = {
    "c1_ref": dict(spec=po.Spec(), P=25, T=16),
    "c2_small": dict(spec=po.Spec(), P=64, T=256),
    "n4_L5": dict(spec=po.Spec(n_layers=5), P=9, T=33),
    "n4_cz_L3": dict(spec=po.Spec(n_layers=3, entangler=po.CZ_RING), P=7, T=40),
    "n4_reup_L3": dict(spec=po.Spec(n_layers=3, reupload=1), P=7, T=40),
    "n3_L4": dict(spec=po.Spec(n_qubits=3, n_layers=4, n_out=2, n_feat=3), P=5, T=17),
    "n2_L3": dict(spec=po.Spec(n_qubits=2, n_layers=3, n_out=2, n_feat=2), P=5, T=17),
    "n6_tiled_reup_L3": dict(spec=po.Spec(n_qubits=6, n_layers=3, n_out=2, n_feat=4, reupload=1, feature_map=po.FEAT_TILED), P=6, T=19),
    "c3_small": dict(spec=po.Spec(n_qubits=8, n_layers=8, n_out=2, n_feat=4, reupload=1, feature_map=po.FEAT_TILED), P=8, T=32),
    "n8_firstk_L2": dict(spec=po.Spec(n_qubits=8, n_layers=2, n_out=2, n_feat=4), P=4, T=16),
    "n10_cz_L2": dict(spec=po.Spec(n_qubits=10, n_layers=2, n_out=2, n_feat=4, entangler=po.CZ_RING, feature_map=po.FEAT_TILED), P=3, T=8),
    "n12_L3": dict(spec=po.Spec(n_qubits=12, n_layers=3, n_out=2, n_feat=4, feature_map=po.FEAT_TILED), P=2, T=6),
}


def golden_inputs(name, cols_sample):
    """Seeded inputs for a golden case (also used by the tests; keep in sync)."""
    g = GOLDEN_SPECS[name]
    spec, P, T = g["spec"], g["P"], g["T"]
    rng = np.random.default_rng(sum(map(ord, name)))
    base = rng.random(spec.n_params).astype(np.float32)
    params = (base[None] + 0.1 * rng.normal(size=(P, spec.n_params))).astype(np.float32)
    rows = rng.choice(cols_sample["sample_obs"].shape[0], T, replace=False)
    nf = spec.n_feat
    obs = np.ascontiguousarray(cols_sample["sample_obs"][rows][:, :nf])
    nxt = np.ascontiguousarray(cols_sample["sample_next_obs"][rows][:, :nf])
    act = cols_sample["sample_act"][rows] % spec.n_out
    rew = cols_sample["sample_rew"][rows].astype(np.float32)
    term = cols_sample["sample_term"][rows].astype(np.float32)
    return spec, params, obs, act, rew, nxt, term


def make_golden_fitness():
    sample = np.load(os.path.join(HERE, "cartpole_v5_sample.npz"))
    out = {}
    for name in GOLDEN_SPECS:
        spec, params, obs, act, rew, nxt, term = golden_inputs(name, sample)
        out[name + "_fitness"] = po.fitness(spec, params, obs, act, rew, nxt, term)
        out[name + "_q"] = np.stack([po.qvalues_tensor(spec, p, obs) for p in params])
        print(name, out[name + "_fitness"][:3])
    np.savez_compressed(os.path.join(HERE, "golden_fitness.npz"), **out)


def make_mini_hdf5(cols):
    """2048 transitions in the (N,1,4)/(N,1,1) layout that dataset.py:27-41 squeezes (the shape of the missing
    offline_cartpole_v2.hdf5 according to the reference's comments), written with tests/golden/hdf5_writer.py."""
    from hdf5_writer import Writer
    n = 2048
    data = {"observations": cols["observations"][:n].reshape(n, 1, 4),
            "next_observations": cols["next_observations"][:n].reshape(n, 1, 4),
            "actions": cols["actions"][:n].reshape(n, 1, 1),
            "rewards": cols["rewards"][:n],
            "terminals": cols["terminals"][:n]}
    chunks = {"observations": (128, 1, 1), "next_observations": (128, 1, 1), "actions": (128, 1, 1),
              "rewards": (128,), "terminals": (512,)}
    Writer().write(os.path.join(HERE, "cartpole_mini_v2layout.hdf5"), data, chunks)


def make_reference_training_run(cols):
    """The reference's OWN train.py / agent.py / optim.py / dataset.py / utils.py / config.py, unmodified, for one
    epoch -- made importable by the three shims in tests/golden/shims (pennylane -> float64 oracle, h5py -> our
    reader, gymnasium -> CartPole restatement).  The data file is the first 10 240 transitions of
    offline_cartpole_test_v5.hdf5 rewritten in the (N,1,4)/(N,1,1) layout train.py's hard-wired file name is
    documented to have; 5 % of it = 512 samples = 32 updates = 2 GA optimisations (agent.py:86)."""
    import contextlib
    import io
    import tempfile
    from hdf5_writer import Writer
    n = 10240
    data = {"observations": cols["observations"][:n].reshape(n, 1, 4),
            "next_observations": cols["next_observations"][:n].reshape(n, 1, 4),
            "actions": cols["actions"][:n].reshape(n, 1, 1), "rewards": cols["rewards"][:n], "terminals": cols["terminals"][:n]}
    chunks = {"observations": (640, 1, 1), "next_observations": (640, 1, 1), "actions": (640, 1, 1), "rewards": (640,), "terminals": (2560,)}
    Writer().write(os.path.join(HERE, "cartpole_train_v2layout.hdf5"), data, chunks)
    out = {}
    for m in [k for k in sys.modules if k in ("optim", "utils", "agent", "train", "dataset", "config")]:
        del sys.modules[m]
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.path.insert(0, os.path.join(HERE, "shims"))
    import train as ref_train
    import agent as ref_agent
    captured = {}
    orig_init = ref_agent.DQNAgent.__init__

    def spy_init(self, *a, **k):           # observation only: keep a handle on the agent run_train builds
        orig_init(self, *a, **k)
        captured["agent"] = self

    ref_agent.DQNAgent.__init__ = spy_init
    ref_train.DQNAgent = ref_agent.DQNAgent
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(HERE, "cartpole_train_v2layout.hdf5"), os.path.join(tmp, "offline_cartpole_v2.hdf5"))
        os.chdir(tmp)
        try:
            for seed in (0, 3):
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    ref_train.run_train("CartPole-v1", 1, seed)
                log = buf.getvalue()
                ag = captured["agent"]
                rewards = [float(l.split("Reward = ")[1]) for l in log.splitlines() if l.startswith("Evaluation Episode")]
                tag = f"seed{seed}_"
                out[tag + "policy_params"] = ag.policy_net.params.detach().cpu().numpy().copy()
                out[tag + "population"] = np.stack([i[0].cpu().numpy() for i in ag.optimizer.population])
                out[tag + "interaction_count"] = np.array(ag.optimizer.interaction_count)
                out[tag + "update_counter"] = np.array(ag.update_counter)
                out[tag + "memory_len"] = np.array(len(ag.memory))
                out[tag + "eval_rewards"] = np.array(rewards)
                out[tag + "rng_probe"] = np.array(ag.rng.random())
                print(f"reference train.py seed {seed}: rewards mean {np.mean(rewards):.2f}, interactions {ag.optimizer.interaction_count}")
        finally:
            os.chdir(cwd)
            torch.use_deterministic_algorithms(False)
    np.savez_compressed(os.path.join(HERE, "train_run_ref.npz"), **out)


def make_es_trajectory():
    """ESOptimizer (no reference counterpart) driven by the float64 ORACLE as its fitness
    backend: the trajectory the CUDA path has to reproduce (tests/test_gpu_parity.py::test_es_optimizer_on_gpu).
    Offspring arithmetic is float32 torch on either side, so identical selections give identical parents."""
    import oqrl_b200
    from oqrl_b200 import ESOptimizer, VQC
    sample = np.load(os.path.join(HERE, "cartpole_v5_sample.npz"))
    rows = np.arange(512)
    cols = [sample["sample_" + k][rows] for k in ("obs", "act", "rew", "next_obs", "term")]

    class Batch:            # an IndexBatch look-alike whose rows the oracle reads from the committed sample
        indices = rows.astype(np.int32)
        dataset = None

        def __len__(self):
            return len(rows)

    def oracle_fitness(spec, params, dataset, idx, gamma=0.99):
        ospec = po.Spec(spec.n_qubits, spec.n_layers, spec.n_out, spec.n_feat, spec.entangler, spec.reupload, spec.feature_map)
        return torch.tensor(po.fitness(ospec, params.cpu().numpy(), *cols, gamma))

    keep = oqrl_b200.engine.population_fitness
    oqrl_b200.engine.population_fitness = oracle_fitness
    try:
        rng = np.random.default_rng(9)
        policy = VQC(4, 2, n_layers=2, rng=rng)
        es = ESOptimizer(policy, mu=16, lam=240, num_generations=8, rng=rng)
        es.parents = es.parents.cpu()
        init = es.parents.numpy().copy()
        es.optimize(None, Batch())
    finally:
        oqrl_b200.engine.population_fitness = keep
    np.savez_compressed(os.path.join(HERE, "es_trajectory_ref.npz"), init_parents=init, final_parents=es.parents.numpy(),
                        last_fitness=np.array(es.last_fitness), sigma=np.array(es.sigma),
                        best=es.best_individual[0].numpy(), rng_probe=np.array(rng.random()))
    print("ES trajectory: best fitness", es.last_fitness[0], "sigma", es.sigma)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "es":
        make_es_trajectory()
        sys.exit(0)
    cols = make_sample()
    make_mini_hdf5(cols)
    make_reference_training_run(cols)
    make_kat()
    make_ga_trajectory(cols)
    make_golden_fitness()
    make_es_trajectory()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
