"""Population sharding across GPUs.

Candidates are independent (optim.py:204), so the population is cut into
contiguous blocks, one per rank; the dataset and the generation's transition
indices are replicated.  The only exchange is one all-gather of the fitness
vector per generation (NCCL over NVLink on GPUs; gloo in the CPU tests), after
which every rank runs the same seeded selection.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_cand: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank`; the first n_cand % world ranks get one extra."""
    base, extra = divmod(n_cand, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedFitness:
    """fitness[P] of the whole population, each rank evaluating only its block.

    ``local_eval(params_block) -> fitness_block`` is the per-rank evaluator
    (normally a closure over engine.population_fitness); it is injectable so the
    exchange logic can be tested with gloo on CPU.
    """

    def __init__(self, n_cand: int, local_eval, group=None, device=None):
        self.n_cand = n_cand
        self.local_eval = local_eval
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(n_cand, self.world, self.rank)
        self.block = -(-n_cand // self.world)          # padded block length used on the wire
        self.device = device
        self._send = None
        self._recv = None

    def local_slice(self, params_full):
        return params_full[self.lo:self.hi]

    def __call__(self, params_block) -> torch.Tensor:
        fit = self.local_eval(params_block)
        if self.world == 1:
            return fit
        dev = fit.device
        if self._send is None or self._send.device != dev:
            self._send = torch.zeros(self.block, dtype=torch.float32, device=dev)
            self._recv = torch.empty(self.block * self.world, dtype=torch.float32, device=dev)
        self._send[: self.hi - self.lo].copy_(fit)
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        if self.n_cand == self.block * self.world:
            return self._recv
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.n_cand, self.world, r)
            parts.append(self._recv[r * self.block: r * self.block + (hi - lo)])
        return torch.cat(parts)
