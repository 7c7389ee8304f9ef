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


class PeerScatterFitness:
    """Same contract as ShardedFitness, but the all-gather is fused into the fitness kernel: every rank's
    full fitness vector lives in symmetric (peer-mapped) memory and each rank's kernel stores its block
    directly into all of them over NVLink while it computes (``oqrl_fitness_scatter``).  One cross-rank
    barrier on peer-mapped flag vectors replaces the NCCL collective -- ``oqrl_peer_barrier``, or, when selection
    follows on the device, the flag wait inside ``oqrl_ga_rank_peers`` (``scatter`` + ``rank_args``).  GPU + NCCL group only.

    ``n_total`` (optional) shards a population that does not divide evenly: rank r evaluates ``shard_bounds(n_total,
    world, r)``; without it every rank owns ``n_cand_per_rank`` candidates.
    """

    def __init__(self, n_cand_per_rank: int, spec, dataset, group=None, own_barrier: bool = True, n_total: int = None):
        import torch.distributed._symmetric_memory as symm_mem
        self.spec, self.dataset = spec, dataset
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if n_total is None:
            self.block = n_cand_per_rank
            self.n_total = self.block * self.world
            self.lo, self.hi = self.rank * self.block, (self.rank + 1) * self.block
        else:
            self.n_total = int(n_total)
            self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
            self.block = self.hi - self.lo
        dev = torch.device("cuda", torch.cuda.current_device())
        # Two generations of buffers: generation g+1 is written into the other buffer, so the only cross-rank
        # synchronisation per generation is "all peers finished writing" (a rank that is one generation ahead
        # can never touch the buffer a slower rank is still reading, because to get two ahead it has to pass
        # the slower rank's barrier of the generation in between).
        self.bufs, self.handles, self.peer_ptrs = [], [], []
        for _ in range(2):
            b = symm_mem.empty(self.n_total, dtype=torch.float32, device=dev)
            h = symm_mem.rendezvous(b, self.group)
            self.bufs.append(b)
            self.handles.append(h)
            self.peer_ptrs.append([int(p) for p in h.buffer_ptrs])
        # flag vectors of the library's own barrier (oqrl_peer_barrier): world + 1 words per rank, epochs only grow
        self.flags = symm_mem.empty(self.world + 1, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.flag_handle = symm_mem.rendezvous(self.flags, self.group)
        self.flag_ptrs = [int(p) for p in self.flag_handle.buffer_ptrs]
        self.own_barrier = own_barrier
        self._gen = 0
        torch.cuda.synchronize()
        self.handles[0].barrier(channel=0)     # every rank's flags are zero before anyone signals
        torch.cuda.synchronize()

    def scatter(self, params_block, idx) -> torch.Tensor:
        """Launch this rank's evaluation with the peer stores, WITHOUT closing the generation: the returned buffer
        is complete only after a barrier with ``rank_args()`` / ``oqrl_peer_barrier`` of the same epoch."""
        from . import engine
        k = self._gen & 1
        self._gen += 1
        engine.population_fitness_scatter(self.spec, params_block, self.dataset, idx, self.peer_ptrs[k], self.lo)
        return self.bufs[k]

    def rank_args(self):
        """(flag_ptrs, rank, epoch) of the generation ``scatter`` just opened -- for ``engine.ga_rank(..., peers=...)``."""
        return self.flag_ptrs, self.rank, self._gen

    def __call__(self, params_block, idx) -> torch.Tensor:
        from . import engine
        buf = self.scatter(params_block, idx)
        if self.own_barrier:
            # launched so that its launch latency hides under the tail of the fitness kernel
            engine.peer_barrier(self.flag_ptrs, self.rank, self._gen)
        else:
            self.handles[(self._gen - 1) & 1].barrier(channel=0)
        return buf                             # every peer's stores into my vector have landed (stream order)

    def host_step(self, h_params, h_idx, h_out=None):
        """One generation end to end from HOST buffers through ``oqrl_fitness_scatter_host``: pinned parameters of this
        rank's block and row indices in; the exchanged vector copied into ``h_out`` on the rank that passes one."""
        from . import engine
        k = self._gen & 1
        self._gen += 1
        engine.population_fitness_scatter_host(self.spec, h_params, self.dataset, h_idx, self.peer_ptrs[k], self.lo,
                                               self.flag_ptrs, self.rank, self._gen, h_out, self.n_total)
        return h_out

    def timed_out(self) -> bool:
        """True if a peer failed to reach a barrier within the kernel's time limit (host sync)."""
        return bool(self.flags[self.world].item())
