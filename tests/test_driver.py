"""Training-driver mirror (oqrl_b200/driver.py) against tests/golden/train_run_ref.npz, which was produced by the
reference's OWN train.py / agent.py / optim.py / dataset.py / utils.py run unmodified behind the shims in
tests/golden/shims (CPU only: the CUDA entry points are replaced by the float64 oracle; this test is about control
flow, replay-memory sampling, RNG order and the evaluation rollouts)."""
import os

import numpy as np
import pytest
import torch

import oqrl_b200
from oracle import pqc_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class FakeDeviceDataset:
    def __init__(self, obs, nxt, act, rew, term):
        self.cols = (np.asarray(obs, np.float32), np.asarray(nxt, np.float32), np.asarray(act).reshape(-1), np.asarray(rew), np.asarray(term))
        self.n_rows, self.n_feat = len(obs), 4

    def __len__(self):
        return self.n_rows


@pytest.fixture
def oracle_engine(monkeypatch):
    def fitness(spec, params, ds, idx, gamma=0.99):
        idx = np.asarray(idx)
        o, n, a, r, t = ds.cols
        p = params.cpu().numpy() if isinstance(params, torch.Tensor) else np.asarray(params)
        return torch.tensor(po.fitness(po.Spec(), p.reshape(-1, 24), o[idx], a[idx], r[idx], n[idx], t[idx], gamma))

    def qvalues(spec, params, x):
        return torch.tensor(po.qvalues_tensor(po.Spec(), params.cpu().numpy().reshape(-1), x.cpu().numpy()))[None]

    monkeypatch.setattr(oqrl_b200.engine, "DeviceDataset", FakeDeviceDataset)
    monkeypatch.setattr(oqrl_b200.engine, "population_fitness", fitness)
    monkeypatch.setattr(oqrl_b200.engine, "qvalues", qvalues)


@pytest.mark.parametrize("seed", [0, 3])
def test_driver_reproduces_reference_training_run(seed, oracle_engine):
    from oqrl_b200.driver import run_train
    ref = dict(np.load(os.path.join(GOLDEN, "train_run_ref.npz")))
    st = run_train(os.path.join(GOLDEN, "cartpole_train_v2layout.hdf5"), 1, seed)
    tag = f"seed{seed}_"
    np.testing.assert_array_equal(st["policy_params"], ref[tag + "policy_params"])
    np.testing.assert_array_equal(np.stack([i[0].numpy() for i in st["agent"].optimizer.population]), ref[tag + "population"])
    assert st["interaction_count"] == int(ref[tag + "interaction_count"]) == 2 * 10 * 25 * 16
    assert st["update_counter"] == int(ref[tag + "update_counter"]) == 32
    assert st["memory_len"] == int(ref[tag + "memory_len"]) == 512
    np.testing.assert_array_equal(np.array(st["eval_rewards"][0]), ref[tag + "eval_rewards"])
    assert st["agent"].rng.random() == float(ref[tag + "rng_probe"])


def test_cartpole_matches_oracle_env():
    from oqrl_b200.cartpole import CartPole
    from oracle.cartpole_env import CartPoleV1
    a, b = CartPole(), CartPoleV1()
    rng = np.random.default_rng(0)
    for seed in (0, 7):
        sa, sb = a.reset(seed=seed)[0], b.reset(seed=seed)[0]
        np.testing.assert_array_equal(sa, sb)
        for _ in range(600):
            act = int(rng.integers(2))
            ra, rb = a.step(act), b.step(act)
            np.testing.assert_array_equal(ra[0], rb[0])
            assert ra[1:4] == rb[1:4]
            if ra[2] or ra[3]:
                break
