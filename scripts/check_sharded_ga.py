"""Multi-GPU correctness of the sharded GA (ShardedDeviceGAOptimizer: oqrl_fitness_scatter + oqrl_ga_rank_peers +
oqrl_ga_breed on every rank).

Run with  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_sharded_ga.py
1. the reference trajectory: the seeded run of tests/golden/ga_trajectory_ref.npz (produced by the reference's own
   optim.py) must come out bit for bit on every rank, although each rank evaluates only its shard of the 25 candidates;
2. a large population (P = 4096 + 3, uneven shards) against the single-GPU DeviceGAOptimizer on the same seed.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oqrl_b200 import DeviceGAOptimizer, GAOptimizer, ShardedDeviceGAOptimizer, VQC, engine  # noqa: E402
from oqrl_b200.dataset import IndexBatch  # noqa: E402


class Exp:
    def __init__(self, obs, action, reward, next_obs, terminated):
        self.obs, self.action, self.reward, self.next_obs, self.terminated = obs, action, reward, next_obs, terminated


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    sample = dict(np.load(os.path.join(ROOT, "tests", "golden", "cartpole_v5_sample.npz")))
    ref = dict(np.load(os.path.join(ROOT, "tests", "golden", "ga_trajectory_ref.npz")))
    bad = 0
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        policy = VQC(4, 2, n_layers=2, rng=rng)
        VQC(4, 2, n_layers=2, rng=rng)
        GAOptimizer(model=policy, rng=rng)
        ga = ShardedDeviceGAOptimizer(policy, rng=rng)
        sel = rng.choice(4096, 16, replace=False)
        batch = [Exp(torch.tensor(sample["sample_obs"][i]), torch.tensor(int(sample["sample_act"][i])), torch.tensor(float(sample["sample_rew"][i])),
                     torch.tensor(sample["sample_next_obs"][i]), torch.tensor(float(sample["sample_term"][i]))) for i in sel]
        ga.optimize(None, batch)
        tag = f"seed{seed}_"
        pop = np.stack([i[0].cpu().numpy() for i in ga.population])
        bad += int(not np.array_equal(pop, ref[tag + "final_population"]))
        bad += int(not np.array_equal(policy.params.cpu().numpy(), ref[tag + "final_model_params"]))
        bad += int(ga.interaction_count != int(ref[tag + "interaction_count"]))
        bad += int(rng.random() != float(ref[tag + "rng_probe"]))
        bad += int(ga._exchange.timed_out())
    # large, unevenly sharded population against the single-GPU optimizer
    ds = engine.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    batch = IndexBatch(ds, np.arange(512, dtype=np.int32))
    pops = []
    for cls in (DeviceGAOptimizer, ShardedDeviceGAOptimizer):
        rng = np.random.default_rng(11)
        policy = VQC(4, 2, n_layers=2, rng=rng)
        ga = cls(policy, population_size=4099, num_generations=4, rng=rng)
        ga.optimize(None, batch)
        pops.append(ga._pop.clone())
    bad += int(not torch.equal(pops[0], pops[1]))
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("sharded GA check:", "OK" if int(t.item()) == 0 else f"MISMATCH ({int(t.item())})", f"world={world}")
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
