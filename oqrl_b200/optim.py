"""Drop-in ``BaseOptimizer`` / ``GAOptimizer`` (reference: /root/reference/src/optim.py).

The GA keeps the reference's interface, side effects and RNG draw order
(SURVEY.md Appendix D); the one change is that a generation's fitness values
come from ONE fused kernel launch over the whole population (optim.py:204)
instead of population x 2 x batch sequential QNode calls.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
import torch

from . import engine

# /root/reference/src/config.py:7-10
POPULATION_SIZE = 25
NUM_GENERATIONS = 10
MUTATION_RATE = 0.1
CROSSOVER_RATE = 0.5
GAMMA_FITNESS = 0.99  # literal at optim.py:163
ELITES = 2            # optim.py:212
PARENT_POOL = 5       # optim.py:220


class BaseOptimizer(ABC):
    """Pluggable-optimiser seam (optim.py:12-23)."""

    def __init__(self, model, rng=None):
        self.model = model
        self.rng = rng or np.random.default_rng()

    @abstractmethod
    def optimize(self, loss_fn, batch):
        """Run the full optimisation loop."""

    @abstractmethod
    def step(self, loss_fn, batch):
        """Optional per-batch step."""


def _column(batch, name, dtype):
    vals = [getattr(e, name) for e in batch]
    vals = [v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=dtype) for v in vals]
    return torch.stack(vals)


def stack_batch(batch):
    """list[Experience] -> (obs, act, rew, next_obs, term) tensors (optim.py:73-112)."""
    obs = _column(batch, "obs", torch.float32)
    act = _column(batch, "action", torch.int64).reshape(-1)     # optim.py:160 .view(-1, 1)
    rew = _column(batch, "reward", torch.float32).reshape(-1)
    nxt = _column(batch, "next_obs", torch.float32)
    term = _column(batch, "terminated", torch.float32).reshape(-1)
    return obs, act, rew, nxt, term


class GAOptimizer(BaseOptimizer):
    def __init__(self, model, population_size=POPULATION_SIZE, num_generations=NUM_GENERATIONS,
                 mutation_rate=MUTATION_RATE, crossover_rate=CROSSOVER_RATE, rng=None):
        super().__init__(model, rng)
        self.population_size = population_size
        self.num_generations = num_generations
        self.mutation_rate = mutation_rate   # stored, never used -- as in the reference (optim.py:39 vs :194)
        self.crossover_rate = crossover_rate
        self.rng = rng or np.random.default_rng()
        self.population = [self._initialize_individual() for _ in range(population_size)]
        self.best_individual = None
        self.interaction_count = 0
        self.last_fitness = None   # fitness list of the most recent generation (host floats)

    # -- population bookkeeping (host side, RNG order of optim.py:51-61,175-197) ----------
    def _initialize_individual(self):
        out = []
        for param in self.model.parameters():
            noise = self.rng.normal(loc=0.0, scale=1.0, size=param.shape)
            noise = torch.tensor(noise, dtype=torch.float32, device=param.device)
            out.append(param.data.clone() + 0.1 * noise)
        return out

    def _crossover(self, parent1, parent2):
        child1, child2 = [], []
        for a, b in zip(parent1, parent2):
            take_a = torch.tensor(self.rng.random(a.shape) < self.crossover_rate, dtype=torch.bool, device=a.device)
            child1.append(torch.where(take_a, a, b))
            child2.append(torch.where(take_a, b, a))
        return child1, child2

    def _mutate(self, individual):
        with torch.no_grad():
            for gene in individual:
                # float64 noise added in place into the float32 genes (optim.py:193-197)
                gene.add_(torch.tensor(self.rng.normal(0, 0.1, size=gene.shape), device=gene.device))

    def _load_into_model(self, individual):
        for param, value in zip(self.model.parameters(), individual):
            param.data.copy_(value)

    # -- fitness -------------------------------------------------------------------------
    def _flat_population(self, individuals):
        return torch.stack([torch.cat([g.reshape(-1) for g in ind]) for ind in individuals])

    def _population_fitness(self, individuals, batch):
        """One launch: fitness of every individual on ``batch`` -> list of python floats."""
        spec = self.model.spec
        flat = self._flat_population(individuals)
        if hasattr(batch, "dataset") and hasattr(batch, "indices"):   # IndexBatch fast path
            fit = engine.population_fitness(spec, flat, batch.dataset, batch.indices, GAMMA_FITNESS)
        else:
            obs, act, rew, nxt, term = stack_batch(batch)
            fit = engine.population_fitness_raw(spec, flat, obs, act, rew, nxt, term, GAMMA_FITNESS)
        return [float(v) for v in fit.to(torch.float64).cpu().tolist()]

    def _evaluate_fitness(self, individual, loss_fn, batch):
        """Fitness of one individual (optim.py:64-173); ``loss_fn`` is ignored as in the reference."""
        self._load_into_model(individual)                 # optim.py:69-70 side effect
        value = self._population_fitness([individual], batch)[0]
        self.interaction_count += len(batch)              # optim.py:171
        return value

    def evaluate_population(self, batch):
        """Batched equivalent of ``[self._evaluate_fitness(ind, ...) for ind in self.population]``."""
        fitness = self._population_fitness(self.population, batch)
        if self.population:
            self._load_into_model(self.population[-1])    # state the sequential loop leaves behind
        self.interaction_count += len(self.population) * len(batch)
        return fitness

    # -- GA loop (optim.py:199-239) ---------------------------------------------------------
    def optimize(self, loss_fn, batch):
        for _generation in range(self.num_generations):
            fitness = self.evaluate_population(batch)
            self.last_fitness = fitness
            order = np.argsort(fitness)[::-1]
            self.population = [self.population[i] for i in order]
            self.best_individual = self.population[0]
            nxt = self.population[:ELITES]
            while len(nxt) < self.population_size:
                pa, pb = self.rng.choice(PARENT_POOL, size=2, replace=False)
                c1, c2 = self._crossover(self.population[pa], self.population[pb])
                self._mutate(c1)
                self._mutate(c2)
                nxt.extend([c1, c2])
            self.population = nxt[: self.population_size]
        if self.best_individual is not None:
            self._load_into_model(self.best_individual)
        else:
            print("Warning: best_individual is None. Skipping weight update.")

    def step(self, loss_fn, batch):
        # the GA has no per-batch step (optim.py:241-243)
        pass


class DeviceGAOptimizer(GAOptimizer):
    """GA with the population resident on the device as one [P, D] tensor.

    Same constructor, RNG draw order and arithmetic as ``GAOptimizer`` -- a seeded run produces bit-identical
    populations -- but a generation is three launches (fused fitness kernel, ``oqrl_ga_rank``, ``oqrl_ga_breed``)
    with no host round trip, instead of ~6 torch calls per child.  ``population`` stays available as a list of
    row views.
    """

    def __init__(self, model, *args, exact_rng_stream=True, **kwargs):
        super().__init__(model, *args, **kwargs)
        # exact_rng_stream=False generates the randomness in the breeding kernel (oqrl_ga_breed_random): same
        # distributions, different stream (no trajectory parity with the reference), no host work per generation
        self.exact_rng_stream = exact_rng_stream
        dev = self._population_device()
        self._pop = torch.stack([torch.cat([g.reshape(-1) for g in ind]) for ind in self.population]).to(dev, torch.float32).contiguous()
        self._shapes = [tuple(p.shape) for p in model.parameters()]

    # -- the device operations of a generation, overridable (sharded subclass; CPU stand-ins in the gloo tests) -------
    def _population_device(self):
        return engine._require_cuda()

    def _rank(self, fit, k):
        return engine.ga_rank(fit, k)

    def _breed(self, order, pair, mask, noise):
        return engine.ga_breed(self._pop, order, ELITES, pair, mask, noise, PARENT_POOL)

    # list-of-lists view expected by callers of the reference class (optim.py:46)
    def _as_individual(self, row):
        out, off = [], 0
        for shp in self._shapes:
            n = int(np.prod(shp)) if shp else 1
            out.append(row[off:off + n].reshape(shp))
            off += n
        return out

    @property
    def population(self):
        if getattr(self, "_pop", None) is None:
            return self._list_population
        return [self._as_individual(self._pop[i]) for i in range(self._pop.shape[0])]

    @population.setter
    def population(self, value):
        self._list_population = value
        if getattr(self, "_pop", None) is not None:
            self._pop = torch.stack([torch.cat([g.reshape(-1) for g in ind]) for ind in value]).to(self._pop.device, torch.float32).contiguous()

    def draw_generation(self, n_genes):
        """The reference's RNG draws of one generation, in its order (optim.py:218-230)."""
        n_pairs = max(0, -(-(self.population_size - ELITES) // 2))
        pair = np.empty((n_pairs, 2), dtype=np.int32)
        mask = np.empty((n_pairs, n_genes), dtype=np.uint8)
        noise = np.empty((n_pairs, 2, n_genes), dtype=np.float64)
        shapes = self._shapes
        if not self.exact_rng_stream:
            # two distinct parents out of the pool: first uniform, second uniform over the remaining ones
            first = self.rng.integers(0, PARENT_POOL, size=n_pairs)
            second = self.rng.integers(0, PARENT_POOL - 1, size=n_pairs)
            pair[:, 0] = first
            pair[:, 1] = second + (second >= first)
            mask[:] = self.rng.random((n_pairs, n_genes)) < self.crossover_rate
            noise[:] = self.rng.normal(0, 0.1, size=(n_pairs, 2, n_genes))
            return pair, mask, noise
        for j in range(n_pairs):
            pair[j] = self.rng.choice(PARENT_POOL, size=2, replace=False)
            off = 0
            for shp in shapes:                                   # _crossover draws per parameter tensor
                n = int(np.prod(shp)) if shp else 1
                mask[j, off:off + n] = (self.rng.random(shp) < self.crossover_rate).reshape(-1)
                off += n
            for child in range(2):                               # _mutate(child1) then _mutate(child2)
                off = 0
                for shp in shapes:
                    n = int(np.prod(shp)) if shp else 1
                    noise[j, child, off:off + n] = self.rng.normal(0, 0.1, size=shp).reshape(-1)
                    off += n
        return pair, mask, noise

    def _prepare_batch(self, batch):
        """Device-resident form of a batch, built once per optimize() call (the same batch serves every generation)."""
        dev = self._pop.device
        if hasattr(batch, "dataset") and hasattr(batch, "indices"):
            return ("idx", batch.dataset, engine._dev_i32(batch.indices, dev))
        obs, act, rew, nxt, term = stack_batch(batch)
        if act.numel() and (int(act.min()) < 0 or int(act.max()) >= self.model.spec.n_out):   # host tensors: no device sync
            raise IndexError("action index outside 0..n_out-1")
        return ("raw", engine._dev_f32(obs, dev), engine._dev_i32(act, dev), engine._dev_f32(rew, dev),
                engine._dev_f32(nxt, dev), engine._dev_f32(term, dev))

    def _fitness_prepared(self, prepared):
        spec = self.model.spec
        if prepared[0] == "idx":
            return engine.population_fitness(spec, self._pop, prepared[1], prepared[2], GAMMA_FITNESS)
        _, obs, act, rew, nxt, term = prepared
        return engine.population_fitness_raw(spec, self._pop, obs, act, rew, nxt, term, GAMMA_FITNESS, validate=False)

    def _fitness_tensor(self, batch):
        return self._fitness_prepared(self._prepare_batch(batch))

    def evaluate_population(self, batch):
        fit = self._fitness_tensor(batch)
        self.interaction_count += self._pop.shape[0] * len(batch)
        return [float(v) for v in fit.to(torch.float64).cpu().tolist()]

    @property
    def last_fitness(self):
        """Fitness list of the most recent generation (host floats; read back on demand)."""
        dev_fit = getattr(self, "_last_fitness_dev", None)
        if dev_fit is not None:
            return [float(v) for v in dev_fit.to(torch.float64).cpu().tolist()]
        return getattr(self, "_last_fitness_host", None)

    @last_fitness.setter
    def last_fitness(self, value):
        self._last_fitness_host = value
        self._last_fitness_dev = None

    def optimize(self, loss_fn, batch):
        """All generations enqueued without a host round trip: fused fitness kernel, selection on the device
        (``oqrl_ga_rank``), one breeding launch -- per generation.  The reference's random draws do not depend on
        the fitness values, so the exact-stream mode draws every generation's randomness up front, in the
        reference's order (optim.py:218-230), and uploads it once."""
        G = self.num_generations
        if G <= 0 or self._pop.shape[0] == 0:
            print("Warning: best_individual is None. Skipping weight update.")
            return
        P, D = self._pop.shape
        if P < PARENT_POOL:
            # the reference draws parent ranks from range(5) and indexes the population with them (optim.py:220-222):
            # IndexError as soon as a drawn rank reaches past a smaller population
            raise IndexError(f"population_size={P} is smaller than the parent pool of {PARENT_POOL} (optim.py:220)")
        dev = self._pop.device
        prepared = self._prepare_batch(batch)
        k = max(ELITES, PARENT_POOL)
        if self.exact_rng_stream:
            draws = [self.draw_generation(D) for _ in range(G)]
            pair = torch.as_tensor(np.stack([d[0] for d in draws])).to(dev)
            mask = torch.as_tensor(np.stack([d[1] for d in draws])).to(dev)
            noise = torch.as_tensor(np.stack([d[2] for d in draws])).to(dev)
        elif getattr(self, "_device_seed", None) is None:
            self._device_seed = int(self.rng.integers(0, 2 ** 63 - 1))      # one draw, at first use
            self._device_generation = 0
        best = None
        fit = None
        if prepared[0] == "idx" and type(self)._rank is DeviceGAOptimizer._rank and type(self)._breed is DeviceGAOptimizer._breed:
            # dataset-resident batch, stock device operations: the whole call is ONE library call (oqrl_ga_run) -- the
            # same kernels in the same order, without ~25 us of interpreter / binding time per launch
            if self.exact_rng_stream:
                self._pop, fit, best = engine.ga_run(self.model.spec, self._pop, prepared[1], prepared[2], G, ELITES, PARENT_POOL,
                                                     draws=(pair.contiguous(), mask.contiguous(), noise.contiguous()), gamma=GAMMA_FITNESS)
            else:
                self._pop, fit, best = engine.ga_run(self.model.spec, self._pop, prepared[1], prepared[2], G, ELITES, PARENT_POOL,
                                                     crossover_rate=self.crossover_rate, sigma=0.1, seed=self._device_seed,
                                                     generation0=self._device_generation, gamma=GAMMA_FITNESS)
                self._device_generation += G
            G_loop = 0
        else:
            G_loop = G
        for g in range(G_loop):
            fit = self._fitness_prepared(prepared)
            order = self._rank(fit, k)
            if g == G - 1:
                best = self._pop.index_select(0, order[:1].to(torch.int64)).reshape(-1)
            if self.exact_rng_stream:
                self._pop = self._breed(order, pair[g], mask[g], noise[g])
            else:
                self._pop = engine.ga_breed_random(self._pop, order, ELITES, PARENT_POOL, self.crossover_rate, 0.1,
                                                   self._device_seed, self._device_generation)
                self._device_generation += 1
        self.interaction_count += G * P * len(batch)
        self._last_fitness_dev = fit
        self.best_individual = self._as_individual(best)
        self._load_into_model(self.best_individual)


class ShardedDeviceGAOptimizer(DeviceGAOptimizer):
    """``DeviceGAOptimizer`` with the population's fitness evaluation sharded over the GPUs of a process group
    (optim.py:204-209 across devices: "all-gather of the fitness vector feeding selection").

    Every rank holds the whole ``[P, D]`` population and the same seeded generator, so breeding needs no exchange:
    per generation a rank evaluates only its contiguous block (``oqrl_fitness_scatter``: the kernel stores each
    fitness value straight into every rank's full vector over NVLink), then ``oqrl_ga_rank_peers`` closes the
    generation (cross-rank flag wait inside the selection kernel -- no barrier launch) and ranks the exchanged
    vector, and ``oqrl_ga_breed`` builds the next population -- identically on every rank.  The trajectory is bit
    for bit the single-GPU ``DeviceGAOptimizer``'s (hence the reference's): a fitness value does not depend on how
    the population is sharded.  ``exact_rng_stream=False`` is not offered: in-kernel randomness is keyed per rank.

    ``batch``: an ``IndexBatch`` over a dataset resident on every rank, or a ``list[Experience]`` (uploaded as a
    temporary table).  Needs an initialised NCCL process group with one rank per GPU.
    """

    def __init__(self, model, *args, group=None, **kwargs):
        kwargs.setdefault("exact_rng_stream", True)
        if not kwargs["exact_rng_stream"]:
            raise ValueError("the sharded GA keeps the reference's RNG stream (identical draws on every rank)")
        super().__init__(model, *args, **kwargs)
        self.group = group
        self._exchange = None
        self._exchange_key = None

    def _make_exchange(self, dataset):
        from .sharding import PeerScatterFitness
        return PeerScatterFitness(0, self.model.spec, dataset, group=self.group, n_total=self._pop.shape[0])

    def _prepare_batch(self, batch):
        dev = self._pop.device
        if hasattr(batch, "dataset") and hasattr(batch, "indices"):
            dataset, idx = batch.dataset, engine._dev_i32(batch.indices, dev)
            engine._check_host_indices(batch.indices, dataset)
        else:
            obs, act, rew, nxt, term = stack_batch(batch)
            if act.numel() and (int(act.min()) < 0 or int(act.max()) >= self.model.spec.n_out):
                raise IndexError("action index outside 0..n_out-1")
            dataset = engine.DeviceDataset(obs.cpu().numpy(), nxt.cpu().numpy(), act.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy())
            idx = torch.arange(len(batch), dtype=torch.int32, device=dev)
        key = (id(dataset), self._pop.shape[0])
        if self._exchange is None or self._exchange_key != key:
            self._exchange, self._exchange_key = self._make_exchange(dataset), key
        return ("sharded", dataset, idx)

    def _fitness_prepared(self, prepared):
        ex = self._exchange
        return ex.scatter(self._pop[ex.lo:ex.hi], prepared[2])      # complete once _rank has waited for the peers

    def _rank(self, fit, k):
        return engine.ga_rank(fit, k, peers=self._exchange.rank_args())

    def optimize(self, loss_fn, batch):
        super().optimize(loss_fn, batch)
        if getattr(self, "_last_fitness_dev", None) is not None:
            self._last_fitness_dev = self._last_fitness_dev.clone()     # the exchange buffer is reused two generations on

    def evaluate_population(self, batch):
        prepared = self._prepare_batch(batch)
        fit = self._exchange(self._pop[self._exchange.lo:self._exchange.hi], prepared[2])    # scatter + barrier
        self.interaction_count += self._pop.shape[0] * len(batch)
        return [float(v) for v in fit.to(torch.float64).cpu().tolist()]


class ESOptimizer(BaseOptimizer):
    """(mu + lambda) evolution strategy behind the same ``BaseOptimizer`` seam (SURVEY.md 8(f) row 4:
    "further optimizers", /root/reference/src/config.py:12, src/train.py:36).  The reference ships no second
    metaheuristic, so there is nothing to mirror; this class shows that the batched fitness entry points are
    optimizer-agnostic: a generation is ONE fitness launch over mu parents + lambda offspring, truncation selection,
    and Gaussian mutation with the 1/5th success rule -- all tensor operations on the population's device.

    Interface as ``GAOptimizer``: ``optimize(loss_fn, batch)`` leaves the best weights in ``model``;
    ``best_individual``, ``interaction_count``, ``last_fitness``; ``batch`` is a ``list[Experience]`` or an
    ``IndexBatch``.
    """

    def __init__(self, model, mu=5, lam=20, num_generations=NUM_GENERATIONS, sigma=0.1, rng=None):
        super().__init__(model, rng)
        self.mu, self.lam, self.num_generations, self.sigma = int(mu), int(lam), int(num_generations), float(sigma)
        self._shapes = [tuple(p.shape) for p in model.parameters()]
        flat = torch.cat([p.data.reshape(-1) for p in model.parameters()]).to(torch.float32)
        noise = torch.as_tensor(self.rng.normal(0.0, 0.1, size=(self.mu, flat.numel())), dtype=torch.float32, device=flat.device)
        self.parents = (flat[None, :] + noise).contiguous()
        if torch.cuda.is_available():                    # keep the population where the fitness kernel runs
            self.parents = self.parents.cuda()
        self.best_individual = None
        self.interaction_count = 0
        self.last_fitness = None

    def _as_individual(self, row):
        out, off = [], 0
        for shp in self._shapes:
            n = int(np.prod(shp)) if shp else 1
            out.append(row[off:off + n].reshape(shp))
            off += n
        return out

    def _fitness(self, pop, batch):
        spec = self.model.spec
        if hasattr(batch, "dataset") and hasattr(batch, "indices"):
            return engine.population_fitness(spec, pop, batch.dataset, batch.indices, GAMMA_FITNESS)
        obs, act, rew, nxt, term = stack_batch(batch)
        return engine.population_fitness_raw(spec, pop, obs, act, rew, nxt, term, GAMMA_FITNESS)

    @staticmethod
    def _select(fit, k):
        """Indices of the k best, best first -- the selection of ``oqrl_ga_rank`` (a stable ascending sort read
        backwards: ties to the higher index).  On the device it IS ``oqrl_ga_rank``; the host branch serves
        fitness vectors that do not live on a GPU (the oracle-backed CPU tests)."""
        if fit.is_cuda:
            return engine.ga_rank(fit.to(torch.float32).contiguous(), k).to(torch.int64)
        order = np.argsort(fit.detach().cpu().numpy(), kind="stable")[::-1][:k]
        return torch.as_tensor(order.copy(), dtype=torch.int64)

    def optimize(self, loss_fn, batch):
        dev = self.parents.device
        D = self.parents.shape[1]
        for _generation in range(self.num_generations):
            pick = torch.as_tensor(self.rng.integers(0, self.mu, size=self.lam), device=dev)
            step = torch.as_tensor(self.rng.normal(0.0, 1.0, size=(self.lam, D)), dtype=torch.float32, device=dev)
            offspring = self.parents[pick] + self.sigma * step
            pop = torch.cat([self.parents, offspring]).contiguous()
            fit = self._fitness(pop, batch)
            self.interaction_count += pop.shape[0] * len(batch)
            order = self._select(fit, self.mu)
            fit = fit.to(pop.device, torch.float64)
            # 1/5th rule on the share of offspring that beat the best parent
            success = float((fit[self.mu:] > fit[:self.mu].max()).to(torch.float64).mean())
            self.sigma *= 1.22 if success > 0.2 else 0.82
            self.parents = pop[order].contiguous()
            self.last_fitness = [float(v) for v in fit[order].cpu().tolist()]
        self.best_individual = self._as_individual(self.parents[0].clone())
        for param, value in zip(self.model.parameters(), self.best_individual):
            param.data.copy_(value)

    def step(self, loss_fn, batch):
        pass
