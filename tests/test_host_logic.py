"""Host-side mirrors (VQC / GAOptimizer) against the reference's own optim.py (CPU only).

tests/golden/ga_trajectory_ref.npz was produced by running the UNMODIFIED
/root/reference/src/optim.py; here the mirror is driven with the same seeds
and the CUDA entry point replaced by the float64 oracle (the oracle is the
checker -- this test is about control flow, side effects and RNG order)."""
import numpy as np
import pytest
import torch

import oqrl_b200
from oqrl_b200 import GAOptimizer, VQC, BaseOptimizer
from oracle import pqc_oracle as po


class Exp:  # utils.py:25-34 stand-in
    def __init__(self, obs, action, reward, next_obs, terminated):
        self.obs, self.action, self.reward, self.next_obs, self.terminated = obs, action, reward, next_obs, terminated


def make_batch(sample_cols, rows):
    return [Exp(torch.tensor(sample_cols["obs"][r]), torch.tensor(int(sample_cols["act"][r]), dtype=torch.int64),
                torch.tensor(float(sample_cols["rew"][r]), dtype=torch.float32), torch.tensor(sample_cols["next_obs"][r]),
                torch.tensor(float(sample_cols["term"][r]), dtype=torch.float32)) for r in rows]


@pytest.fixture
def oracle_engine(monkeypatch):
    def fake_raw(spec, params, obs, act, rew, next_obs, term, gamma=0.99):
        ospec = po.Spec(spec.n_qubits, spec.n_layers, spec.n_out, spec.n_feat, spec.entangler, spec.reupload, spec.feature_map)
        f = po.fitness(ospec, params.cpu().numpy(), obs.cpu().numpy(), act.cpu().numpy(), rew.cpu().numpy(),
                       next_obs.cpu().numpy(), term.cpu().numpy(), gamma)
        return torch.tensor(f)   # float64: same ordering as the reference's python floats
    monkeypatch.setattr(oqrl_b200.engine, "population_fitness_raw", fake_raw)


def sample_cols(sample):
    return {k: sample["sample_" + k] for k in ("obs", "act", "rew", "next_obs", "term")}


@pytest.mark.parametrize("seed", [0, 1])
def test_ga_trajectory_matches_reference_optim(seed, ga_ref, sample, oracle_engine):
    cols = sample_cols(sample)
    rng = np.random.default_rng(seed)
    policy = VQC(4, 2, n_layers=2, rng=rng)
    VQC(4, 2, n_layers=2, rng=rng)                         # target net draw (agent.py:90)
    GAOptimizer(model=policy, rng=rng)                     # agent.py:99
    ga = GAOptimizer(policy, rng=rng)                      # train.py:37
    pos = rng.choice(4096, 16, replace=False)               # positions inside the committed sample
    rows = pos
    tag = f"seed{seed}_"
    np.testing.assert_array_equal(sample["sample_idx"][pos], ga_ref[tag + "rows"])
    np.testing.assert_array_equal(np.stack([i[0].numpy() for i in ga.population]), ga_ref[tag + "init_population"])
    batch = make_batch(cols, rows)
    log = []
    orig = ga.evaluate_population
    ga.evaluate_population = lambda b: (log.append(orig(b)) or log[-1])
    ga.optimize(None, batch)
    np.testing.assert_allclose(np.array(log), ga_ref[tag + "fitness_log"], rtol=1e-12)
    np.testing.assert_array_equal(np.stack([i[0].numpy() for i in ga.population]), ga_ref[tag + "final_population"])
    np.testing.assert_array_equal(policy.params.numpy(), ga_ref[tag + "final_model_params"])
    np.testing.assert_array_equal(ga.best_individual[0].numpy(), ga_ref[tag + "best_individual"])
    assert ga.interaction_count == int(ga_ref[tag + "interaction_count"]) == 10 * 25 * 16
    assert rng.random() == float(ga_ref[tag + "rng_probe"])      # RNG stream position identical


def test_single_candidate_side_effects(sample, oracle_engine):
    rng = np.random.default_rng(3)
    policy = VQC(4, 2, rng=rng)
    ga = GAOptimizer(policy, population_size=5, num_generations=1, rng=rng)
    cols = {k: sample["head_" + k] for k in ("obs", "act", "rew", "next_obs", "term")}
    batch = make_batch(cols, range(16))
    v = ga._evaluate_fitness(ga.population[3], None, batch)
    assert isinstance(v, float) and ga.interaction_count == 16          # optim.py:171,173
    np.testing.assert_array_equal(policy.params.numpy(), ga.population[3][0].numpy())   # optim.py:69-70
    want = po.fitness(po.Spec(), ga.population[3][0].numpy()[None], cols["obs"][:16], cols["act"][:16], cols["rew"][:16],
                      cols["next_obs"][:16], cols["term"][:16])[0]
    assert v == pytest.approx(want, rel=1e-12)
    elite_before = None
    ga.optimize(None, batch)
    elite_before = ga.population[0]
    assert ga.best_individual is elite_before                           # elites carried by reference (optim.py:209,212)
    assert len(ga.population) == 5 and ga.step(None, batch) is None


def test_vqc_init_and_interfaces():
    rng = np.random.default_rng(0)
    m = VQC(4, 2, n_layers=2, rng=rng)
    np.testing.assert_array_equal(m.params.numpy(), np.random.default_rng(0).random(24).astype(np.float32))
    assert m.params.requires_grad is False and list(m.state_dict()) == ["params"]
    m2 = VQC(4, 2, rng=rng)
    m2.load_state_dict(m.state_dict())
    np.testing.assert_array_equal(m2.params.numpy(), m.params.numpy())
    assert (m.input_dim, m.output_dim, m.n_layers) == (4, 2, 2)
    with pytest.raises(TypeError):
        BaseOptimizer(m)                                                # abstract (optim.py:12-23)
    big = VQC(4, 2, n_layers=8, rng=rng, n_qubits=8, reupload=True, feature_map=oqrl_b200.FEAT_TILED)
    assert big.params.numel() == 8 * 8 * 3 and big.spec.n_feat == 4


def test_es_optimizer_behind_the_same_seam(sample, oracle_engine):
    """a second metaheuristic on the batched fitness entry point (oracle-backed here).
    Plus-selection never loses the best individual, the run is seed-deterministic, and the model ends up with it."""
    from oqrl_b200 import ESOptimizer
    cols = sample_cols(sample)
    batch = make_batch(cols, range(16))
    runs = []
    for _ in range(2):
        rng = np.random.default_rng(5)
        policy = VQC(4, 2, n_layers=2, rng=rng)
        es = ESOptimizer(policy, mu=4, lam=12, num_generations=6, rng=rng)
        first = max(float(v) for v in oqrl_b200.engine.population_fitness_raw(policy.spec, es.parents, *[t for t in __import__("oqrl_b200").optim.stack_batch(batch)]))
        es.optimize(None, batch)
        assert es.last_fitness[0] >= first - 1e-12 and es.last_fitness == sorted(es.last_fitness, reverse=True)
        assert es.interaction_count == 6 * 16 * 16
        np.testing.assert_array_equal(policy.params.detach().cpu().numpy().reshape(-1), es.best_individual[0].cpu().numpy().reshape(-1))
        runs.append(es.parents.cpu().numpy().copy())
    np.testing.assert_array_equal(runs[0], runs[1])
