"""Thin torch-facing wrappers over the C ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and the stream; every circuit
evaluation happens inside liboqrl_b200.so.  There is no CPU path -- without a
CUDA device these functions raise.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import OqrlError, check, lib, probe
from .spec import CircuitSpec


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("oqrl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev_f32(t, dev) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t, dtype=np.float32))
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _dev_i32(t, dev) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=dev, dtype=torch.int32).contiguous()


class DeviceDataset:
    """HBM-resident transition table (``oqrl_dataset``): uploaded once, indexed per generation."""

    def __init__(self, obs, next_obs, act, rew, term):
        dev = _require_cuda()
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        next_obs = np.ascontiguousarray(next_obs, dtype=np.float32)
        if obs.ndim != 2 or obs.shape != next_obs.shape:
            raise ValueError("obs/next_obs must be [N, n_feat] with equal shapes")
        n = obs.shape[0]
        act = np.ascontiguousarray(np.asarray(act).reshape(n), dtype=np.int32)
        rew = np.ascontiguousarray(np.asarray(rew).reshape(n), dtype=np.float32)
        term = np.ascontiguousarray(np.asarray(term).reshape(n), dtype=np.float32)
        self.n_rows, self.n_feat = int(n), int(obs.shape[1])
        self.device = dev
        self.h2d_bytes = obs.nbytes + next_obs.nbytes + act.nbytes + rew.nbytes + term.nbytes
        handle = ctypes.c_void_p()
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        check(lib().oqrl_dataset_upload(vp(obs), vp(next_obs), vp(act), vp(rew), vp(term),
                                        self.n_rows, self.n_feat, _stream(), ctypes.byref(handle)))
        self._h = handle

    def __len__(self):
        return self.n_rows

    def close(self):
        if getattr(self, "_h", None):
            lib().oqrl_dataset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def gather(self, idx):
        """Rows ``idx`` -> (obs, next_obs, act, rew, term) device tensors."""
        idx = _dev_i32(idx, self.device)
        n = idx.numel()
        if n and (int(idx.min()) < 0 or int(idx.max()) >= self.n_rows):
            raise IndexError("dataset index out of range")
        obs = torch.empty((n, self.n_feat), dtype=torch.float32, device=self.device)
        nxt = torch.empty_like(obs)
        act = torch.empty(n, dtype=torch.int32, device=self.device)
        rew = torch.empty(n, dtype=torch.float32, device=self.device)
        term = torch.empty(n, dtype=torch.float32, device=self.device)
        check(lib().oqrl_dataset_gather(self._h, idx.data_ptr(), n, obs.data_ptr(), nxt.data_ptr(),
                                        act.data_ptr(), rew.data_ptr(), term.data_ptr(), _stream()))
        return obs, nxt, act, rew, term


def _check_host_indices(idx, dataset) -> None:
    """IndexError for a row outside the table, as the reference's fancy indexing raises (dataset.py:60-66).  Only
    indices that are still on the host are checked (``IndexBatch`` keeps a numpy array, so this costs no device
    sync); indices already on the device are range-checked by the kernels, which write NaN fitness instead."""
    if isinstance(idx, torch.Tensor):
        if idx.is_cuda or idx.numel() == 0:
            return
        lo, hi = int(idx.min()), int(idx.max())
    else:
        a = np.asarray(idx)
        if a.size == 0:
            return
        lo, hi = int(a.min()), int(a.max())
    if lo < 0 or hi >= dataset.n_rows:
        raise IndexError(f"dataset index out of range: {lo}..{hi} outside 0..{dataset.n_rows - 1}")


def _population_matrix(params: torch.Tensor, spec: CircuitSpec) -> torch.Tensor:
    """[P, n_params]; a layer-free circuit (n_params == 0) keeps its candidate count from the leading dimension."""
    if spec.n_params == 0:
        return params.reshape(params.shape[0] if params.dim() > 1 else 1, 0)
    return params.reshape(-1, spec.n_params)


def qvalues(spec: CircuitSpec, params, x) -> torch.Tensor:
    """<Z_i> for every (candidate, row): params [P,D] (or [D]), x [B,n_feat] -> float32 [P,B,n_out]."""
    dev = _require_cuda()
    params = _dev_f32(params, dev)
    params = params.reshape(1, -1) if params.ndim == 1 else params
    if params.ndim != 2 or params.shape[1] != spec.n_params:
        raise ValueError(f"params must be [P, {spec.n_params}]")
    x = _dev_f32(x, dev)
    x = x.reshape(1, -1) if x.ndim == 1 else x
    if x.ndim != 2 or x.shape[1] != spec.n_feat:
        raise ValueError(f"x must be [B, {spec.n_feat}]")
    P, B = params.shape[0], x.shape[0]
    q = torch.empty((P, B, spec.n_out), dtype=torch.float32, device=dev)
    cs = spec.to_c()
    check(lib().oqrl_qvalues(ctypes.byref(cs), params.data_ptr(), P, x.data_ptr(), B, q.data_ptr(), _stream()))
    return q


def population_fitness(spec: CircuitSpec, params, dataset: DeviceDataset, idx, gamma: float = 0.99) -> torch.Tensor:
    """fitness[P] (float32, device) of a population on dataset rows ``idx`` (optim.py:204)."""
    dev = _require_cuda()
    params = _population_matrix(_dev_f32(params, dev), spec)
    _check_host_indices(idx, dataset)
    idx = _dev_i32(idx, dev)
    P, T = params.shape[0], idx.numel()
    out = torch.empty(P, dtype=torch.float32, device=dev)
    cs = spec.to_c()
    check(lib().oqrl_fitness(ctypes.byref(cs), params.data_ptr(), P, dataset._h, idx.data_ptr(), T,
                             float(gamma), out.data_ptr(), _stream()))
    return out


def population_fitness_scatter(spec: CircuitSpec, params, dataset: DeviceDataset, idx, peer_ptrs, cand_offset: int,
                               gamma: float = 0.99) -> None:
    """Fitness of the local candidate block written by the kernel straight into every peer's full fitness
    vector (``peer_ptrs`` = device pointers of those vectors as mapped in this process).  See
    ``oqrl_fitness_scatter``; the caller barriers across ranks before reading."""
    dev = _require_cuda()
    params = _population_matrix(_dev_f32(params, dev), spec)
    _check_host_indices(idx, dataset)
    idx = _dev_i32(idx, dev)
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
    cs = spec.to_c()
    check(lib().oqrl_fitness_scatter(ctypes.byref(cs), params.data_ptr(), params.shape[0], dataset._h, idx.data_ptr(),
                                     idx.numel(), float(gamma), arr, len(peer_ptrs), int(cand_offset), _stream()))


def population_fitness_scatter_host(spec: CircuitSpec, h_params: torch.Tensor, dataset: DeviceDataset, h_idx: torch.Tensor,
                                    peer_ptrs, cand_offset: int, flag_ptrs, rank: int, epoch: int, h_out, n_total: int,
                                    gamma: float = 0.99) -> None:
    """Sharded evaluation with HOST buffers (``oqrl_fitness_scatter_host``): this rank's pinned parameter block and the
    row indices in, the exchange closed by the library's barrier, and the whole exchanged vector copied into ``h_out``
    (a pinned float32 tensor of ``n_total`` values) on the rank that passes one -- ``None`` elsewhere.  Synchronises."""
    _require_cuda()
    assert h_params.dtype == torch.float32 and h_idx.dtype == torch.int32 and not h_params.is_cuda and not h_idx.is_cuda
    assert h_out is None or (h_out.dtype == torch.float32 and not h_out.is_cuda and h_out.numel() >= n_total)
    P = h_params.shape[0]
    fp = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
    fl = (ctypes.c_void_p * len(flag_ptrs))(*[ctypes.c_void_p(int(p)) for p in flag_ptrs])
    cs = spec.to_c()
    check(lib().oqrl_fitness_scatter_host(ctypes.byref(cs), h_params.data_ptr(), P, dataset._h, h_idx.data_ptr(), h_idx.numel(),
                                          float(gamma), fp, len(peer_ptrs), int(cand_offset), fl, int(rank), int(epoch) & 0xFFFFFFFF,
                                          h_out.data_ptr() if h_out is not None else None, int(n_total), _stream()))


def peer_barrier(flag_ptrs, rank: int, epoch: int) -> None:
    """Cross-rank barrier on peer-mapped flag vectors (``oqrl_peer_barrier``), asynchronous on the current stream."""
    _require_cuda()
    arr = (ctypes.c_void_p * len(flag_ptrs))(*[ctypes.c_void_p(int(p)) for p in flag_ptrs])
    check(lib().oqrl_peer_barrier(arr, len(flag_ptrs), int(rank), int(epoch) & 0xFFFFFFFF, _stream()))


def ga_breed(pop: torch.Tensor, order, n_elite: int, pair, mask, noise, parent_pool: int = 5) -> torch.Tensor:
    """Next population [P, D] in one launch (``oqrl_ga_breed``): elites + crossover children + mutation noise.
    pop float32 [P, D] on the device; order int32 [>= max(n_elite, parent_pool)]; pair int32 [n_pairs, 2] ranks
    below ``parent_pool``; mask uint8 [n_pairs, D]; noise float64 [n_pairs, 2, D] (host arrays are uploaded)."""
    dev = _require_cuda()
    assert pop.is_cuda and pop.dtype == torch.float32 and pop.is_contiguous()
    P, D = pop.shape
    order = _dev_i32(order, dev)
    pair = _dev_i32(pair, dev).reshape(-1, 2)
    mask = torch.as_tensor(np.asarray(mask, dtype=np.uint8) if not isinstance(mask, torch.Tensor) else mask).to(dev, torch.uint8).contiguous()
    noise = torch.as_tensor(np.asarray(noise, dtype=np.float64) if not isinstance(noise, torch.Tensor) else noise).to(dev, torch.float64).contiguous()
    n_pairs = pair.shape[0]
    out = torch.empty_like(pop)
    check(lib().oqrl_ga_breed(pop.data_ptr(), P, D, order.data_ptr(), order.numel(), int(n_elite), pair.data_ptr(),
                              int(parent_pool), mask.data_ptr(), noise.data_ptr(), n_pairs, out.data_ptr(), _stream()))
    return out


def ga_rank(fitness: torch.Tensor, k: int, peers=None) -> torch.Tensor:
    """int32 [k] on the device: indices of the k largest fitness values, best first (``oqrl_ga_rank``).
    ``peers = (flag_ptrs, rank, epoch)`` selects ``oqrl_ga_rank_peers``: the kernel first closes the generation of a
    sharded evaluation (cross-rank flag exchange) and then ranks the exchanged vector."""
    _require_cuda()
    assert fitness.is_cuda and fitness.dtype == torch.float32 and fitness.is_contiguous()
    order = torch.empty(k, dtype=torch.int32, device=fitness.device)
    if peers is None:
        check(lib().oqrl_ga_rank(fitness.data_ptr(), fitness.numel(), int(k), order.data_ptr(), _stream()))
    else:
        flag_ptrs, rank, epoch = peers
        arr = (ctypes.c_void_p * len(flag_ptrs))(*[ctypes.c_void_p(int(p)) for p in flag_ptrs])
        check(lib().oqrl_ga_rank_peers(fitness.data_ptr(), fitness.numel(), int(k), order.data_ptr(), arr, len(flag_ptrs),
                                       int(rank), int(epoch) & 0xFFFFFFFF, _stream()))
    return order


def ga_breed_random(pop: torch.Tensor, order: torch.Tensor, n_elite: int, parent_pool: int, crossover_rate: float,
                    sigma: float, seed: int, generation: int) -> torch.Tensor:
    """Next population with the randomness generated in the kernel (``oqrl_ga_breed_random``)."""
    _require_cuda()
    assert pop.is_cuda and pop.dtype == torch.float32 and pop.is_contiguous()
    assert order.is_cuda and order.dtype == torch.int32 and order.numel() >= max(n_elite, parent_pool)
    out = torch.empty_like(pop)
    check(lib().oqrl_ga_breed_random(pop.data_ptr(), pop.shape[0], pop.shape[1], order.data_ptr(), int(n_elite),
                                     int(parent_pool), float(crossover_rate), float(sigma),
                                     int(seed) & 0xFFFFFFFFFFFFFFFF, int(generation) & 0xFFFFFFFF, out.data_ptr(), _stream()))
    return out


def ga_run(spec: CircuitSpec, pop: torch.Tensor, dataset: DeviceDataset, idx: torch.Tensor, n_generations: int, n_elite: int,
           parent_pool: int, draws=None, crossover_rate: float = 0.5, sigma: float = 0.1, seed: int = 0, generation0: int = 0,
           gamma: float = 0.99):
    """A whole ``optimize()`` call enqueued by one library call (``oqrl_ga_run``).  ``draws`` = (pair [G,n_pairs,2] int32,
    mask [G,n_pairs,D] uint8, noise [G,n_pairs,2,D] float64) device tensors for the reference's RNG stream, or None for
    in-kernel randomness.  Returns (next population, last generation's fitness, that generation's best individual)."""
    dev = _require_cuda()
    assert pop.is_cuda and pop.dtype == torch.float32 and pop.is_contiguous() and pop.shape[1] == spec.n_params
    _check_host_indices(idx, dataset)
    d_idx = _dev_i32(idx, dev)
    other = torch.empty_like(pop)
    fit = torch.empty(pop.shape[0], dtype=torch.float32, device=dev)
    k = max(int(n_elite), int(parent_pool))
    order = torch.empty(k, dtype=torch.int32, device=dev)
    best = torch.empty(pop.shape[1], dtype=torch.float32, device=dev)
    in_b = ctypes.c_int32(0)
    if draws is not None:
        pair, mask, noise = draws
        assert pair.is_cuda and pair.dtype == torch.int32 and pair.is_contiguous() and pair.shape[0] == n_generations
        assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
        assert noise.is_cuda and noise.dtype == torch.float64 and noise.is_contiguous()
        n_pairs, ptrs = pair.shape[1], (pair.data_ptr(), mask.data_ptr(), noise.data_ptr())
    else:
        n_pairs, ptrs = 0, (None, None, None)
    cs = spec.to_c()
    check(lib().oqrl_ga_run(ctypes.byref(cs), pop.data_ptr(), other.data_ptr(), pop.shape[0], dataset._h, d_idx.data_ptr(),
                            d_idx.numel(), float(gamma), int(n_generations), int(n_elite), int(parent_pool), ptrs[0], ptrs[1], ptrs[2],
                            int(n_pairs), float(crossover_rate), float(sigma), int(seed) & 0xFFFFFFFFFFFFFFFF,
                            int(generation0) & 0xFFFFFFFF, fit.data_ptr(), order.data_ptr(), best.data_ptr(), ctypes.byref(in_b), _stream()))
    return (other if in_b.value else pop), fit, best


def population_fitness_raw(spec: CircuitSpec, params, obs, act, rew, next_obs, term, gamma: float = 0.99,
                           validate: bool = True) -> torch.Tensor:
    """Same as population_fitness for an explicit batch of transitions (device or host arrays).
    ``validate=False`` skips the action-range check (a device read-back) for batches checked before."""
    dev = _require_cuda()
    params = _population_matrix(_dev_f32(params, dev), spec)
    obs = _dev_f32(obs, dev).reshape(-1, spec.n_feat)
    next_obs = _dev_f32(next_obs, dev).reshape(-1, spec.n_feat)
    T = obs.shape[0]
    act = _dev_i32(act, dev).reshape(-1)
    rew = _dev_f32(rew, dev).reshape(-1)
    term = _dev_f32(term, dev).reshape(-1)
    if not (next_obs.shape[0] == act.numel() == rew.numel() == term.numel() == T):
        raise ValueError("transition columns disagree on the batch size")
    if validate and T and (int(act.min()) < 0 or int(act.max()) >= spec.n_out):
        raise IndexError("action index outside 0..n_out-1")   # torch.gather raises in the reference
    P = params.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=dev)
    cs = spec.to_c()
    check(lib().oqrl_fitness_raw(ctypes.byref(cs), params.data_ptr(), P, obs.data_ptr(), next_obs.data_ptr(),
                                 act.data_ptr(), rew.data_ptr(), term.data_ptr(), T, float(gamma),
                                 out.data_ptr(), _stream()))
    return out


def population_fitness_host(spec: CircuitSpec, h_params: torch.Tensor, dataset: DeviceDataset,
                            h_idx: torch.Tensor, h_out: torch.Tensor, gamma: float = 0.99) -> torch.Tensor:
    """End-to-end call with HOST (pinned) buffers: H2D params+indices, kernel, D2H fitness, sync."""
    _require_cuda()
    assert h_params.dtype == torch.float32 and h_idx.dtype == torch.int32 and h_out.dtype == torch.float32
    assert not h_params.is_cuda and not h_idx.is_cuda and not h_out.is_cuda
    P = h_params.shape[0]
    cs = spec.to_c()
    check(lib().oqrl_fitness_host(ctypes.byref(cs), h_params.data_ptr(), P, dataset._h, h_idx.data_ptr(),
                                  h_idx.numel(), float(gamma), h_out.data_ptr(), _stream()))
    return h_out


def statevector(spec: CircuitSpec, params, x) -> torch.Tensor:
    """Final state of one circuit, complex64 [2^n], logical order (wire 0 = MSB)."""
    dev = _require_cuda()
    params = _dev_f32(params, dev).reshape(-1)
    x = _dev_f32(x, dev).reshape(-1)
    out = torch.empty((1 << spec.n_qubits, 2), dtype=torch.float32, device=dev)
    cs = spec.to_c()
    check(lib().oqrl_statevector(ctypes.byref(cs), params.data_ptr(), x.data_ptr(), out.data_ptr(), _stream()))
    return torch.view_as_complex(out)


def kernel_class(spec: CircuitSpec) -> int:
    cs = spec.to_c()
    rc = lib().oqrl_kernel_class(ctypes.byref(cs))
    if rc < 0:
        check(rc)
    return rc


def kernel_name(spec: CircuitSpec) -> str:
    """Name of the fitness kernel ``spec`` dispatches to (as ncu lists it)."""
    cs = spec.to_c()
    name = lib().oqrl_kernel_name(ctypes.byref(cs))
    if name is None:
        raise OqrlError(-1, lib().oqrl_last_error().decode(errors="replace"))
    return name.decode()


def work_model(spec: CircuitSpec):
    """(executed fp32 flop per circuit-eval, HBM state sweeps per circuit-eval)."""
    cs = spec.to_c()
    f, s = ctypes.c_double(), ctypes.c_double()
    check(probe().oqrl_work_model(ctypes.byref(cs), ctypes.byref(f), ctypes.byref(s)))
    return f.value, s.value


def launch_count() -> int:
    return int(lib().oqrl_launch_count())


def fma_peak_tflops(variant: int = 1, iters: int = 1 << 14) -> float:
    _require_cuda()
    t = ctypes.c_double()
    check(probe().oqrl_fma_peak(variant, iters, ctypes.byref(t), _stream()))
    return t.value


def gate_rate_tflops(form: int = 0, warps_per_smsp: int = 2, iters: int = 1 << 12) -> float:
    """Register-only arithmetic rate of the HBM-class gate body (probe library; see oqrl_gate_rate)."""
    _require_cuda()
    t = ctypes.c_double()
    check(probe().oqrl_gate_rate(form, warps_per_smsp, iters, ctypes.byref(t), _stream()))
    return t.value


def exchange_rate(medium: int, iters: int = 200):
    """(microseconds per re-tiling of one 16-qubit state, aggregate GB/s) -- probe library, see oqrl_exchange_rate."""
    _require_cuda()
    us, gbs = ctypes.c_double(), ctypes.c_double()
    check(probe().oqrl_exchange_rate(int(medium), int(iters), ctypes.byref(us), ctypes.byref(gbs), _stream()))
    return us.value, gbs.value
