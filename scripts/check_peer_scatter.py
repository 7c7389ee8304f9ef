"""Multi-GPU correctness of the fused fitness exchange (oqrl_fitness_scatter + oqrl_peer_barrier).

Run with  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_peer_scatter.py
Every rank evaluates its block of a different population per generation, with deliberately skewed ranks, and
compares the gathered vector bit for bit with its own single-GPU evaluation of the whole population.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oqrl_b200 import CircuitSpec, engine  # noqa: E402
from oqrl_b200.sharding import PeerScatterFitness  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rng = np.random.default_rng(7)                      # same stream on every rank
    n_rows, T = 4096, 256
    ds = engine.DeviceDataset(rng.normal(size=(n_rows, 4)).astype(np.float32), rng.normal(size=(n_rows, 4)).astype(np.float32),
                              rng.integers(0, 2, n_rows).astype(np.int32), np.ones(n_rows, np.float32),
                              (rng.random(n_rows) < 0.05).astype(np.float32))
    worst = 0
    for spec, block in ((CircuitSpec(), 700), (CircuitSpec(6, 2, 2, 4, 0, 1, 1), 96)):
        ps = PeerScatterFitness(block, spec, ds)
        for gen in range(12):
            pop = torch.tensor(rng.normal(size=(block * world, spec.n_params)).astype(np.float32), device="cuda")
            idx = torch.tensor(rng.integers(0, n_rows, T).astype(np.int32), device="cuda")
            if (gen + rank) % 3 == 0:
                time.sleep(0.02 * (rank + 1))           # skew the ranks: the barrier has to hold the fast ones
            got = ps(pop[rank * block:(rank + 1) * block], idx).clone()
            want = engine.population_fitness(spec, pop, ds, idx)
            torch.cuda.synchronize()
            bad = int((got != want).sum().item())
            worst = max(worst, bad)
            assert not ps.timed_out(), "peer barrier timed out"
        del ps
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("peer scatter check:", "OK" if int(t.item()) == 0 else f"MISMATCH ({int(t.item())} values)", f"world={world}")
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
