"""Population sharding + fitness all-gather with gloo, world_size 2, on CPU.

The per-rank evaluator is injected (oracle-backed) -- this exercises the partition, the padded
all-gather and the identical-selection-on-every-rank property, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oqrl_b200.sharding import ShardedFitness, shard_bounds


def test_shard_bounds_cover_population():
    for n, w in ((25, 2), (4096, 8), (7, 8), (16384, 4), (1, 3)):
        blocks = [shard_bounds(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_cand, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from oracle import pqc_oracle as po
    sample = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cartpole_v5_sample.npz"))
    rng = np.random.default_rng(0)                                  # same seed on every rank: replicated inputs
    params = (rng.random(24)[None] + 0.1 * rng.normal(size=(n_cand, 24))).astype(np.float32)
    rows = rng.choice(4096, 16, replace=False)
    cols = [sample["sample_" + k][rows] for k in ("obs", "act", "rew", "next_obs", "term")]

    def local_eval(block):
        return torch.tensor(po.fitness(po.Spec(), block.numpy(), *cols), dtype=torch.float32)

    sf = ShardedFitness(n_cand, local_eval)
    assert (sf.lo, sf.hi) == shard_bounds(n_cand, world, rank)
    full = sf(sf.local_slice(torch.tensor(params)))
    again = sf(sf.local_slice(torch.tensor(params)))                # buffers are reused
    assert torch.equal(full, again)
    order = np.argsort(full.numpy())[::-1]                          # optim.py:207 on every rank
    np.save(os.path.join(out_dir, f"fit{rank}.npy"), full.numpy())
    np.save(os.path.join(out_dir, f"order{rank}.npy"), order)
    if rank == 0:
        ref = po.fitness(po.Spec(), params, *cols).astype(np.float32)
        np.save(os.path.join(out_dir, "ref.npy"), ref)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cand", [25, 64])
def test_sharded_fitness_allgather_gloo(tmp_path, n_cand):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_cand, str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / "ref.npy")
    f0, f1 = np.load(tmp_path / "fit0.npy"), np.load(tmp_path / "fit1.npy")
    assert f0.shape == (n_cand,)
    np.testing.assert_array_equal(f0, f1)                           # every rank holds the same full vector
    np.testing.assert_array_equal(f0, ref)                          # and it is the unsharded result
    np.testing.assert_array_equal(np.load(tmp_path / "order0.npy"), np.load(tmp_path / "order1.npy"))


# ---- sharded GA optimizer: host logic with CPU stand-ins for the three device operations -------------------------
def _ga_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from oracle import pqc_oracle as po
    from oqrl_b200 import GAOptimizer, ShardedDeviceGAOptimizer, VQC
    from oqrl_b200.optim import ELITES, stack_batch
    from test_host_logic import Exp
    sample = np.load(os.path.join(here, "golden", "cartpole_v5_sample.npz"))
    ref = np.load(os.path.join(here, "golden", "ga_trajectory_ref.npz"))

    class GlooExchange:
        """all-gather of the oracle-evaluated shard: the contract of PeerScatterFitness.scatter / rank_args"""
        def __init__(self, n_total, cols):
            self.lo, self.hi = shard_bounds(n_total, world, rank)
            self.n_total, self.cols, self.calls = n_total, cols, 0

        def scatter(self, block, idx):
            assert block.shape[0] == self.hi - self.lo
            self.calls += 1
            mine = torch.tensor(po.fitness(po.Spec(), block.numpy(), *self.cols), dtype=torch.float32)
            parts = [torch.empty(shard_bounds(self.n_total, world, r)[1] - shard_bounds(self.n_total, world, r)[0]) for r in range(world)]
            dist.all_gather(parts, mine) if len({p.numel() for p in parts}) == 1 else _uneven_all_gather(parts, mine)
            return torch.cat(parts)

        def rank_args(self):
            return None

        def timed_out(self):
            return False

    def _uneven_all_gather(parts, mine):
        width = max(p.numel() for p in parts)
        padded = [torch.zeros(width) for _ in parts]
        send = torch.zeros(width)
        send[: mine.numel()] = mine
        dist.all_gather(padded, send)
        for p, q in zip(parts, padded):
            p.copy_(q[: p.numel()])

    class CpuShardedGA(ShardedDeviceGAOptimizer):
        def _population_device(self):
            return torch.device("cpu")

        def _prepare_batch(self, batch):
            obs, act, rew, nxt, term = stack_batch(batch)
            if self._exchange is None:
                self._exchange = GlooExchange(self._pop.shape[0], [obs.numpy(), act.numpy(), rew.numpy(), nxt.numpy(), term.numpy()])
            return ("sharded", None, None)

        def _rank(self, fit, k):                      # oqrl_ga_rank's order: stable ascending sort read backwards
            return torch.as_tensor(np.argsort(fit.numpy(), kind="stable")[::-1][:k].copy(), dtype=torch.int32)

        def _breed(self, order, pair, mask, noise):   # numpy statement of ga_breed_kernel (optim.py:175-197,212-232)
            pop, o = self._pop.numpy(), order.numpy()
            rows = [pop[o[i]] for i in range(ELITES)]
            for j in range(pair.shape[0]):
                a, b = pop[o[int(pair[j, 0])]], pop[o[int(pair[j, 1])]]
                m = mask[j].numpy().astype(bool)
                rows.append((np.where(m, a, b).astype(np.float64) + noise[j, 0].numpy()).astype(np.float32))
                rows.append((np.where(m, b, a).astype(np.float64) + noise[j, 1].numpy()).astype(np.float32))
            return torch.tensor(np.stack(rows)[: pop.shape[0]])

    ok = True
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        policy = VQC(4, 2, n_layers=2, rng=rng)
        VQC(4, 2, n_layers=2, rng=rng)
        GAOptimizer(model=policy, rng=rng)
        ga = CpuShardedGA(policy, rng=rng)
        sel = rng.choice(4096, 16, replace=False)
        batch = [Exp(torch.tensor(sample["sample_obs"][i]), torch.tensor(int(sample["sample_act"][i])), torch.tensor(float(sample["sample_rew"][i])),
                     torch.tensor(sample["sample_next_obs"][i]), torch.tensor(float(sample["sample_term"][i]))) for i in sel]
        ga.optimize(None, batch)
        tag = f"seed{seed}_"
        ok &= np.array_equal(np.stack([i[0].numpy() for i in ga.population]), ref[tag + "final_population"])
        ok &= np.array_equal(policy.params.numpy(), ref[tag + "final_model_params"])
        ok &= np.array_equal(ga.best_individual[0].numpy(), ref[tag + "best_individual"])
        ok &= ga.interaction_count == int(ref[tag + "interaction_count"]) and rng.random() == float(ref[tag + "rng_probe"])
        ok &= ga._exchange.calls == 10 and (ga._exchange.hi - ga._exchange.lo) in (12, 13)      # 25 candidates over two ranks
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array(bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ga_optimizer_gloo(tmp_path):
    """ShardedDeviceGAOptimizer's host logic on two gloo ranks: each rank evaluates its shard of the 25 candidates
    (oracle), the vector is all-gathered, and selection + breeding run identically on both ranks -- the reference's
    GA trajectory (tests/golden/ga_trajectory_ref.npz, from the reference's own optim.py) comes out bit for bit."""
    world = 2
    mp.spawn(_ga_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert bool(np.load(tmp_path / "ok0.npy")) and bool(np.load(tmp_path / "ok1.npy"))
