#!/usr/bin/env python
"""Benchmark of the PQC fitness path (BASELINE.json metric: PQC circuit-evals/sec).

A "step" is one generation's fitness evaluation of the whole population on one
batch of transitions = 2 * P * T circuit evaluations (s and s', optim.py:152,162).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1] [--impl reference]

N = 1: BASELINE config 2 (4-qubit, 2-layer CartPole PQC, P = 4096 x T = 1024) unless --workload says otherwise.
N > 1 (launched by torchrun, one rank per GPU): weak scaling -- every rank evaluates its own block of
`P` candidates, the dataset is replicated, and each step ends with the NCCL all-gather of the fitness vector.

One JSON line is printed by rank 0 (see the contract in the task statement): `value` is device-timed with the
inputs resident in HBM; `e2e` goes through the host-buffer C-ABI call (pinned host params + indices in, fitness out);
`roofline` is the dominant kernel against the measured FP32 FMA peak; `cpu_baseline` is the C oracle on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_qubits, n_layers, n_out, n_feat, entangler, reupload, feature_map, P, T, description)
    "c1": (4, 2, 2, 4, 0, 0, 0, 25, 16, "config 1: reference defaults, 4-qubit 2-layer PQC, population 25 x batch 16"),
    "c2": (4, 2, 2, 4, 0, 0, 0, 4096, 1024, "config 2: 4-qubit 2-layer CartPole PQC, population 4096 x 1024 transitions"),
    "c3": (8, 8, 2, 4, 0, 1, 1, 16384, 1024, "config 3: 8-qubit 8-layer data-reuploading PQC (tiled features), population 16384 x 1024 transitions"),
    "c4": (16, 2, 2, 4, 0, 0, 0, 1024, 256, "config 4: 16-qubit 2-layer PQC, population 1024 x 256 transitions"),
    # config 5 is a single forward circuit (no TD pair): P = T = 1, one circuit-eval per step
    "c5": (24, 8, 2, 4, 0, 0, 1, 1, 1, "config 5: one 24-qubit PQC forward pass (HBM-resident 128 MiB state), tiled features"),
}
METRIC = "PQC circuit-evals/sec (population x transitions x 2)"
UNIT = "circuit-evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--population", type=int, default=0, help="override P (per GPU)")
    ap.add_argument("--transitions", type=int, default=0, help="override T")
    ap.add_argument("--layers", type=int, default=0, help="override the layer count (config 5 sweep)")
    ap.add_argument("--qubits", type=int, default=0, help="override the qubit count (config 5 sweep)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU fitness exchange: peer stores fused into the kernel (p2p) or NCCL all-gather")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--entangler", default="", choices=["", "sel", "cz"],
                    help="override the workload's entangler (sel = the reference's CNOT ring, cz = CZ ring)")
    ap.add_argument("--general-kernel", action="store_true",
                    help="A/B aid: run n = 8 on the general 5..13-qubit shared-memory kernel instead of the specialised one")
    return ap.parse_args()


def workload(args):
    n, L, no, nf, ent, reup, fmap, P, T, desc = WORKLOADS[args.workload]
    if args.population:
        P = args.population
    if args.transitions:
        T = args.transitions
    if args.layers:
        L = args.layers
    if args.qubits:
        n = args.qubits
    if args.entangler:
        ent = 0 if args.entangler == "sel" else 1
    return dict(n=n, L=L, n_out=no, n_feat=nf, ent=ent, reup=reup, fmap=fmap, P=P, T=T, desc=desc, forward_only=args.workload == "c5")


def synthetic_inputs(w, rank, n_rows_dataset=4096):
    """Parameters with the reference's distributions (agent.py:30, optim.py:58-60); transitions = committed rows
    of offline_cartpole_test_v5.hdf5 (the real file is not on the GPU box), tiled if T exceeds them."""
    sample = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_v5_sample.npz"))
    rng = np.random.default_rng(1000 + rank)
    D = w["L"] * w["n"] * 3
    base = np.random.default_rng(0).random(D).astype(np.float32)
    params = (base[None] + 0.1 * rng.normal(size=(w["P"], D))).astype(np.float32)
    nf = w["n_feat"]
    cols = dict(obs=np.ascontiguousarray(sample["sample_obs"][:, :nf]), next_obs=np.ascontiguousarray(sample["sample_next_obs"][:, :nf]),
                act=sample["sample_act"] % w["n_out"], rew=sample["sample_rew"].astype(np.float32), term=sample["sample_term"].astype(np.float32))
    idx = np.random.default_rng(0).choice(n_rows_dataset, size=w["T"], replace=w["T"] > n_rows_dataset).astype(np.int32)
    return params, cols, idx


# ----------------------------------------------------------------------------------------------
# CPU legs (the only places the oracle is executed by this file)
# ----------------------------------------------------------------------------------------------
def cpu_port_rate(w, params, cols, idx, seconds, threads=0):
    """C oracle (oracle/pqc_oracle.c, OpenMP over candidates) on a time-bounded slice of the same workload."""
    from oracle import c_oracle, pqc_oracle as po
    spec = po.Spec(w["n"], w["L"], w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])
    threads = threads or len(os.sched_getaffinity(0))
    T = len(idx)
    obs, nxt, act, rew, term = (cols[k][idx] for k in ("obs", "next_obs", "act", "rew", "term"))
    per_eval = max(po.flops_per_eval(spec), 1) / 2.0e9            # crude: ~2 GFLOP/s/core scalar fp64
    slice_p = int(max(threads, min(w["P"], threads * max(1, int(1.0 / max(per_eval * 2 * T, 1e-9))))))
    c_oracle.fitness(spec, params[:threads], obs[: min(T, 8)], act[: min(T, 8)], rew[: min(T, 8)], nxt[: min(T, 8)], term[: min(T, 8)], nthreads=threads)
    done, t0, lo = 0, time.perf_counter(), 0
    while True:
        hi = min(w["P"], lo + slice_p)
        c_oracle.fitness(spec, params[lo:hi], obs, act, rew, nxt, term, nthreads=threads)
        done += 2 * (hi - lo) * T
        lo = hi if hi < w["P"] else 0
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, threads, f"{done} circuit-evals ({done // (2 * T)} candidates x {T} transitions x 2) in {el:.1f} s"


def cpu_reference_loop_rate(w, params, cols, idx, seconds):
    """Reference-structured per-sample Python loop (oracle/ref_port.py), single thread like the reference."""
    from oracle import pqc_oracle as po, ref_port
    spec = po.Spec(w["n"], w["L"], w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])
    T = min(len(idx), 16)                                          # config.py:4 BATCH_SIZE
    obs, nxt, act, rew, term = (cols[k][idx[:T]] for k in ("obs", "next_obs", "act", "rew", "term"))
    done, t0, p = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        ref_port.fitness_per_candidate(spec, params[p:p + 1], obs, act, rew, nxt, term)
        done += 2 * T
        p = (p + 1) % len(params)
    return done / (time.perf_counter() - t0)


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU path.  The reference itself (PennyLane QNode loop) cannot be installed
    offline, so this arm times the oracle port of the same algorithm on all host threads (kind = "port")."""
    if rank != 0:
        return
    params, cols, idx = synthetic_inputs(w, 0)
    if w["forward_only"]:
        base = cpu_forward_baseline(w, params, cols, idx)
        print(json.dumps({"impl": "reference", "metric": "PQC circuit-evals/sec (single forward circuit)", "value": base["value"], "unit": UNIT,
                          "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 / base["value"], "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(w, args, world),
                          "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_port_rate(w, params, cols, idx, 0.5)
    rates, cores, sample = [], 0, ""
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, cores, sample = cpu_port_rate(w, params, cols, idx, per_step)
        rates.append(r)
    total = time.perf_counter() - t0
    value = statistics.median(rates)
    loop_rate = cpu_reference_loop_rate(w, params, cols, idx, 3.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic params (reference distributions) + committed rows of offline_cartpole_test_v5.hdf5",
        "config": config_dict(w, args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step: {sample}; C oracle (fp64, OpenMP over candidates)",
                         "reference_structured_python_loop": {"value": loop_rate, "unit": UNIT, "cores": 1,
                                                              "note": "per-candidate x per-sample loop of optim.py:204/agent.py:51 without PennyLane tape overhead"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "PennyLane (requirements.txt:7) is not installable offline; this arm is the oracle port, not the unmodified reference",
    }
    print(json.dumps(line), flush=True)


def canonical_flops(w):
    """F = 2^n * [6 * n_embedded * (1 + reupload*(L-1)) + 14*n*L + (3 + n_out)]."""
    n_emb = w["n"] if (w["fmap"] == 1 and w["n_feat"] > 0) else min(w["n_feat"], w["n"])
    reps = w["L"] if (w["reup"] and w["L"] > 0) else 1
    return (1 << w["n"]) * (6 * n_emb * reps + 14 * w["n"] * w["L"] + 3 + w["n_out"])


def config_dict(w, args, world):
    return {"workload": w["desc"], "n_qubits": w["n"], "n_layers": w["L"], "population_per_gpu": w["P"], "transitions": w["T"],
            "circuit_evals_per_step": (1 if w["forward_only"] else 2) * w["P"] * w["T"] * world,
            "parallelism": (f"population-sharded dp{world}" if not w["forward_only"] else f"{world} independent replica(s)"),
            "l2": "flushed between timed steps (256 MiB write)", "entangler": "SEL_CNOT" if w["ent"] == 0 else "CZ_RING",
            "reupload": bool(w["reup"])}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    """HBM roofline denominator: MEASURED_PEAKS.json (driver-written) or the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from oqrl_b200 import CircuitSpec, engine
    from oqrl_b200.sharding import ShardedFitness

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback for the product path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = CircuitSpec(w["n"], w["L"], w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"],
                       general_kernel=int(args.general_kernel))
    params_np, cols, idx_np = synthetic_inputs(w, rank)
    P, T = w["P"], w["T"]
    fwd = w["forward_only"]
    kclass = engine.kernel_class(spec)

    # one-time residency: the whole transition table goes to HBM (dataset.py -> oqrl_dataset_upload)
    ds = engine.DeviceDataset(cols["obs"], cols["next_obs"], cols["act"], cols["rew"], cols["term"])
    d_params = torch.tensor(params_np, device=dev)
    d_idx = torch.tensor(idx_np, device=dev)
    d_x = torch.tensor(cols["obs"][idx_np], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    collective = "none"
    if fwd:
        # single-circuit path does not shard: N ranks = N independent replicas (DESIGN.md "replicas only")
        def kernel_call():
            return engine.qvalues(spec, d_params, d_x)
        step = kernel_call
        evals_per_step = 1.0 * P * T * world
    else:
        def kernel_call():
            return engine.population_fitness(spec, d_params, ds, d_idx)
        evaluator = ShardedFitness(P * world, lambda blk: engine.population_fitness(spec, blk, ds, d_idx), device=dev)
        collective = "none" if world == 1 else "nccl"
        p2p = None
        if world > 1 and args.collective in ("auto", "p2p"):
            try:
                from oqrl_b200.sharding import PeerScatterFitness
                p2p = PeerScatterFitness(P, spec, ds)
                p2p(d_params, d_idx)
                torch.cuda.synchronize()
                collective = "p2p (peer stores fused into the fitness kernel + oqrl_peer_barrier)"
            except Exception as e:  # symmetric memory unavailable: keep NCCL
                if args.collective == "p2p":
                    raise
                p2p = None
                collective = f"nccl (p2p unavailable: {type(e).__name__})"

        def step():
            return p2p(d_params, d_idx) if p2p is not None else evaluator(d_params)
        evals_per_step = 2.0 * P * T * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed value ----------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = engine.launch_count()
    barrier()
    for a, b in ev:
        flush.fill_(1)                      # L2 flush, outside the timed region of the step
        a.record()
        step()
        b.record()
    barrier()
    launches = engine.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = evals_per_step / (ms_per_step * 1e-3)

    # ---- kernel-only duration (events directly around the library call, same stream) -----------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        flush.fill_(1)
        a.record()
        kernel_call()
        b.record()
    torch.cuda.synchronize()
    kern_ms = statistics.median([a.elapsed_time(b) for a, b in kev])

    # ---- end to end through the host-buffer path: pinned host inputs in, host result out, every step ----
    h_params = torch.tensor(params_np).pin_memory()
    h_idx = torch.tensor(idx_np).pin_memory()
    h_x = torch.tensor(cols["obs"][idx_np]).pin_memory()
    h_out = torch.empty(P * world if not fwd else P * T * w["n_out"], dtype=torch.float32).pin_memory()
    e2e_path = "torch H2D copies of the pinned inputs, library call, D2H copy of the result, synchronise"
    if fwd:
        h2d, d2h = h_params.numel() * 4 + h_x.numel() * 4, h_out.numel() * 4

        def e2e_step():
            q = engine.qvalues(spec, h_params.to(dev, non_blocking=True), h_x.to(dev, non_blocking=True))
            h_out.copy_(q.reshape(-1), non_blocking=True)
            torch.cuda.synchronize()
    elif world == 1:
        h2d, d2h = h_params.numel() * 4 + h_idx.numel() * 4, P * 4

        e2e_path = ("oqrl_fitness_host per step: pinned host parameters read in place by the kernel over PCIe (once per work item), "
                    "row indices copied H2D, fitness written in place into pinned host memory, stream synchronised"
                    if os.environ.get("OQRL_HOST_ZEROCOPY", "1") != "0" else
                    "oqrl_fitness_host per step: cudaMemcpyAsync H2D of parameters and indices, kernel, cudaMemcpyAsync D2H, stream synchronised")

        def e2e_step():
            engine.population_fitness_host(spec, h_params, ds, h_idx, h_out)     # the C-ABI host call
    else:
        h2d, d2h = h_params.numel() * 4 + h_idx.numel() * 4, P * world * 4

        gather_buf = torch.empty(P * world, dtype=torch.float32, device=dev)

        def e2e_step():
            blk = h_params.to(dev, non_blocking=True)
            ix = h_idx.to(dev, non_blocking=True)
            if p2p is not None:                      # fused exchange: peer stores + oqrl_peer_barrier
                h_out.copy_(p2p(blk, ix), non_blocking=True)
            else:
                fit = engine.population_fitness(spec, blk, ds, ix)
                dist.all_gather_into_tensor(gather_buf, fit)
                h_out.copy_(gather_buf, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = evals_per_step * args.steps / float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    flops_exec, sweeps = engine.work_model(spec)
    flops_canon = canonical_flops(w)
    evals_per_kernel_call = evals_per_step / world
    kernel_name = engine.kernel_name(spec)
    sm_clk = (clocks or {}).get("sm_mhz") or 0.0
    if fwd and kclass == 2:
        state_bytes = (1 << w["n"]) * 8
        alg_bytes = sweeps * 2.0 * state_bytes * evals_per_kernel_call
        peak, src = measured_hbm_peak()
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "kernel": kernel_name, "kernel_ms": kern_ms, "state_sweeps_per_eval": sweeps, "state_bytes": state_bytes,
                    "algorithmic_bytes_per_eval": sweeps * 2.0 * state_bytes,
                    "gate_by_gate_sweeps": w["n"] * (w["L"] + (1 if w["n_feat"] else 0)) + 1,
                    "executed_tflops": flops_exec * evals_per_kernel_call / (kern_ms * 1e-3) / 1e12, "peak_source": src}
    else:
        peak_scalar = max(engine.fma_peak_tflops(0, 1 << 13), engine.fma_peak_tflops(2, 1 << 13))
        peak_packed = max(engine.fma_peak_tflops(1, 1 << 13), engine.fma_peak_tflops(3, 1 << 13))
        peak = max(peak_scalar, peak_packed)
        achieved = flops_exec * evals_per_kernel_call / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                    "kernel": kernel_name, "kernel_ms": kern_ms, "flop_per_eval_executed": flops_exec,
                    "flop_per_eval_canonical": flops_canon,
                    "canonical_equiv_tflops": flops_canon * evals_per_kernel_call / (kern_ms * 1e-3) / 1e12,
                    "peak_source": "FP32 FMA microbenchmark measured in this process (oqrl_fma_peak: scalar FFMA %.1f, packed FFMA2 %.1f TFLOP/s); "
                                   "MEASURED_PEAKS.json has no FP32 figure. Nominal 148 SM x 128 lanes x 2 x f_clk = %.1f TFLOP/s at the sampled %.0f MHz"
                                   % (peak_scalar, peak_packed, 148 * 128 * 2 * sm_clk * 1e6 / 1e12, sm_clk)}

    # DRAM traffic per launch of the default workload from the committed `ncu --set full` capture of this kernel
    # (profiles/r1_reg_fitness_v8_final_ncu.txt: dram__bytes_read.sum 593 920 B, dram__bytes_write.sum 0 B); a number
    # taken under the profiler, so it is quoted, not re-measured here.  Other workloads: null.
    if args.workload == "c2" and world == 1 and not (args.population or args.transitions or args.layers or args.qubits or args.entangler):
        roofline["traffic"] = 593920
        roofline["traffic_source"] = "profiles/r1_reg_fitness_v8_final_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"

    if kclass == 0 and not fwd:
        # same launch with the last-layer light-cone pruning switched off: every Rot gate of every layer executed
        import dataclasses
        spec_np = dataclasses.replace(spec, no_prune=1)
        for _ in range(3):
            engine.population_fitness(spec_np, d_params, ds, d_idx)
        kev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in kev2:
            flush.fill_(1)
            a.record()
            engine.population_fitness(spec_np, d_params, ds, d_idx)
            b.record()
        torch.cuda.synchronize()
        ms2 = statistics.median([a.elapsed_time(b) for a, b in kev2])
        f2, _ = engine.work_model(spec_np)
        ach2 = f2 * evals_per_kernel_call / (ms2 * 1e-3) / 1e12
        roofline["unpruned_variant"] = {"kernel_ms": ms2, "flop_per_eval_executed": f2, "achieved": ach2, "frac": ach2 / roofline["peak"],
                                        "circuit_evals_per_s": evals_per_kernel_call / (ms2 * 1e-3),
                                        "note": "all gates of every layer applied (only |0>+embedding+layer 0 folded into the product state)"}

    line = {"metric": METRIC if not fwd else "PQC circuit-evals/sec (single forward circuit)", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic params (reference distributions) + committed rows of offline_cartpole_test_v5.hdf5",
            "config": config_dict(w, args, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "path": e2e_path},
            "gpu_launches": int(launches), "roofline": roofline}
    line["config"]["collective"] = collective
    if not args.no_cpu_baseline and world == 1:
        if fwd:
            line["cpu_baseline"] = cpu_forward_baseline(w, params_np, cols, idx_np)
        else:
            rate, cores, sample_desc = cpu_port_rate(w, params_np, cols, idx_np, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample_desc + "; C oracle (fp64, OpenMP over candidates)"}
            if w["n"] <= 8:
                loop_rate = cpu_reference_loop_rate(w, params_np, cols, idx_np, 3.0)
                line["cpu_baseline"]["reference_structured_python_loop"] = {
                    "value": loop_rate, "unit": UNIT, "cores": 1,
                    "note": "per-candidate x per-sample loop (optim.py:204, agent.py:51) without PennyLane tape overhead"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_forward_baseline(w, params, cols, idx):
    """Single large circuit on the C oracle (one thread: the port parallelises over circuits, not inside one).
    Bounded: at most 2 layers are simulated and the time is scaled linearly to the configured depth."""
    from oracle import c_oracle, pqc_oracle as po
    Ls = min(w["L"], 2)
    spec = po.Spec(w["n"], Ls, w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])
    D = Ls * w["n"] * 3
    x = cols["obs"][idx[:1]]
    t0 = time.perf_counter()
    c_oracle.qvalues(spec, params[:1, :D], x, nthreads=1)
    el = time.perf_counter() - t0
    scale = (w["L"] / Ls) if Ls else 1.0
    return {"value": 1.0 / (el * scale), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"one {w['n']}-qubit circuit with {Ls} of {w['L']} layers in {el:.2f} s on the C oracle (fp64, 1 thread), scaled x{scale:g}"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    w = workload(args)
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
