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
