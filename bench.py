#!/usr/bin/env python
"""Benchmark of the PQC fitness path (BASELINE.json metric: PQC circuit-evals/sec).

A "step" is one generation's fitness evaluation of the whole population on one
batch of transitions = 2 * P * T circuit evaluations (s and s', optim.py:152,162).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

N = 1: BASELINE config 2 (4-qubit, 2-layer CartPole PQC, P = 4096 x T = 1024) unless --workload says otherwise; the
line also carries an `also` map with short runs of configs 1, 3, 4 and 5 (value, roofline, a parity bit against the
oracle on a candidate sample, a 3-second CPU baseline) so that every configuration of BASELINE.json is driver-visible.
N > 1 (launched by torchrun, one rank per GPU): weak scaling -- every rank evaluates its own block of `P` candidates,
the dataset is replicated, and each step ends with the exchange of the fitness vector (peer stores fused into the
fitness kernel); the line adds `c3_strong` -- config 3 with its population of 16384 sharded over the N ranks, the
configuration BASELINE.json names for sharding -- and `exchange_verified`: rank 0 recomputes the whole population on
its own GPU and compares the gathered vector bit for bit.

One JSON line is printed by rank 0 (see the contract in the task statement): `value` is device-timed with the
inputs resident in HBM; `e2e` goes through the host-buffer C-ABI call (pinned host params + indices in, fitness out);
`roofline` is the dominant kernel against the measured FP32 FMA peak (executed flop: `flop_basis`) or the measured
HBM copy bandwidth; `cpu_baseline` is the C oracle on the host cores, with the reference's own Python files (behind
the pennylane shim) timed beside it.
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
METRIC_FWD = "PQC circuit-evals/sec (single forward circuit)"
UNIT = "circuit-evals/s"
DATA = "synthetic params (reference distributions) + committed rows of offline_cartpole_test_v5.hdf5"
TOL = 1e-5     # north_star: Q-values and fitness within 1e-5 relative in fp32 (floor max(|ref|, 1))


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
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other configurations")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU fitness exchange: peer stores fused into the kernel (p2p) or NCCL all-gather")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--entangler", default="", choices=["", "sel", "cz"],
                    help="override the workload's entangler (sel = the reference's CNOT ring, cz = CZ ring)")
    ap.add_argument("--general-kernel", action="store_true",
                    help="A/B aid: run n = 8 on the general 5..13-qubit shared-memory kernel instead of the specialised one")
    return ap.parse_args()


def workload(args, name=None):
    name = name or args.workload
    n, L, no, nf, ent, reup, fmap, P, T, desc = WORKLOADS[name]
    if name == args.workload:      # overrides apply to the headline workload only
        P = args.population or P
        T = args.transitions or T
        L = args.layers or L
        n = args.qubits or n
        if args.entangler:
            ent = 0 if args.entangler == "sel" else 1
    return dict(name=name, n=n, L=L, n_out=no, n_feat=nf, ent=ent, reup=reup, fmap=fmap, P=P, T=T, desc=desc, forward_only=name == "c5")


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


def oracle_spec(w):
    from oracle import pqc_oracle as po
    return po.Spec(w["n"], w["L"], w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])


def rel_err(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)))


# ----------------------------------------------------------------------------------------------
# CPU legs (the only places the oracle is executed by this file)
# ----------------------------------------------------------------------------------------------
def cpu_port_rate(w, params, cols, idx, seconds, threads=0):
    """C oracle (oracle/pqc_oracle.c, OpenMP over candidates) on a time-bounded slice of the same workload."""
    from oracle import c_oracle, pqc_oracle as po
    spec = oracle_spec(w)
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


def cpu_reference_files_rate(w, params, cols, idx, seconds):
    """The reference's OWN optim.py / agent.py (oracle/_ref/, copied unmodified by oracle/make_ref.py) behind the
    pennylane shim, single thread like the reference -- or None where the copies are absent or the circuit is not
    one the reference can express (it hard-wires StronglyEntanglingLayers + a single embedding, agent.py:44-45)."""
    from oracle import ref_runner
    if not ref_runner.available() or w["ent"] != 0 or w["reup"] or w["fmap"] != 0 or w["n_feat"] != w["n"]:
        return None
    T = min(len(idx), 64)
    obs, nxt, act, rew, term = (cols[k][idx[:T]] for k in ("obs", "next_obs", "act", "rew", "term"))
    rate, n_cand, _ = ref_runner.fitness_rate_subprocess(w["n"], w["L"], w["n_out"], params[:256], obs, act, rew, nxt, term, seconds)
    return {"value": rate, "unit": UNIT, "cores": 1, "kind": "reference+shim",
            "sample": f"{n_cand} candidates x {T} transitions x 2 through GAOptimizer._evaluate_fitness (optim.py:64-173) / VQC.forward "
                      f"(agent.py:48-61), the reference's own files, pennylane replaced by the float64 oracle shim",
            "note": "an upper bound of the real reference: the shim's QNode is far cheaper than PennyLane's tape machinery"}


def cpu_reference_loop_rate(w, params, cols, idx, seconds):
    """Reference-structured per-sample Python loop (oracle/ref_port.py), single thread like the reference."""
    from oracle import ref_port
    spec = oracle_spec(w)
    T = min(len(idx), 16)                                          # config.py:4 BATCH_SIZE
    obs, nxt, act, rew, term = (cols[k][idx[:T]] for k in ("obs", "next_obs", "act", "rew", "term"))
    done, t0, p = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        ref_port.fitness_per_candidate(spec, params[p:p + 1], obs, act, rew, nxt, term)
        done += 2 * T
        p = (p + 1) % len(params)
    return done / (time.perf_counter() - t0)


def cpu_forward_baseline(w, params, cols, idx, want_q=False):
    """Single large circuit on the C oracle (one thread: the port parallelises over circuits, not inside one).
    Bounded: at most 2 layers are simulated and the time is scaled linearly to the configured depth."""
    from oracle import c_oracle, pqc_oracle as po
    Ls = min(w["L"], 2)
    spec = po.Spec(w["n"], Ls, w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])
    D = Ls * w["n"] * 3
    x = cols["obs"][idx[:1]]
    t0 = time.perf_counter()
    q = c_oracle.qvalues(spec, params[:1, :D], x, nthreads=1)
    el = time.perf_counter() - t0
    scale = (w["L"] / Ls) if Ls else 1.0
    out = {"value": 1.0 / (el * scale), "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"one {w['n']}-qubit circuit with {Ls} of {w['L']} layers in {el:.2f} s on the C oracle (fp64, 1 thread), scaled x{scale:g}"}
    return (out, q, Ls) if want_q else out


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU path on the host cores.  `value` = the C/OpenMP oracle port on all host
    threads (the strongest CPU baseline we can build; kind "port"); the reference's own optim.py / agent.py behind the
    pennylane shim are timed beside it (cpu_baseline.reference_files, kind "reference+shim") -- PennyLane itself is
    not installable offline, so the unmodified reference cannot run end to end."""
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""      # the reference's utils.device must be the CPU (and this arm needs no GPU)
    params, cols, idx = synthetic_inputs(w, 0)
    evals_per_step = (1 if w["forward_only"] else 2) * w["P"] * w["T"] * world
    if w["forward_only"]:
        base = cpu_forward_baseline(w, params, cols, idx)
        print(json.dumps({"impl": "reference", "metric": METRIC_FWD, "value": base["value"], "unit": UNIT,
                          "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 / base["value"], "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": DATA, "config": config_dict(w, world),
                          "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_port_rate(w, params, cols, idx, 0.5)
    rates, cores, sample = [], 0, ""
    for _ in range(args.steps):
        r, cores, sample = cpu_port_rate(w, params, cols, idx, per_step)
        rates.append(r)
    value = statistics.median(rates)
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"each step: {sample}; C oracle (fp64, OpenMP over candidates)"}
    ref_files = cpu_reference_files_rate(w, params, cols, idx, 6.0)
    if ref_files:
        base["reference_files"] = ref_files
    base["reference_structured_python_loop"] = {
        "value": cpu_reference_loop_rate(w, params, cols, idx, 3.0), "unit": UNIT, "cores": 1,
        "note": "per-candidate x per-sample loop of optim.py:204/agent.py:51 without PennyLane tape overhead"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup,
        # the time ONE step of the workload (all of its circuit evaluations) takes at the measured rate; a step of this
        # arm runs a time-bounded slice of it (cpu_baseline.sample)
        "ms_per_step": 1e3 * evals_per_step / value, "slice_seconds_per_step": per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": DATA,
        "config": config_dict(w, world), "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "value = C/OpenMP port of the reference path on all host threads; the reference's own Python files run behind a "
                "pennylane shim in cpu_baseline.reference_files (PennyLane, requirements.txt:7, is not installable offline)",
    }
    print(json.dumps(line), flush=True)


def canonical_flops(w):
    """F = 2^n * [6 * n_embedded * (1 + reupload*(L-1)) + 14*n*L + (3 + n_out)]."""
    n_emb = w["n"] if (w["fmap"] == 1 and w["n_feat"] > 0) else min(w["n_feat"], w["n"])
    reps = w["L"] if (w["reup"] and w["L"] > 0) else 1
    return (1 << w["n"]) * (6 * n_emb * reps + 14 * w["n"] * w["L"] + 3 + w["n_out"])


def config_dict(w, world):
    """Identical in both arms (the driver compares it)."""
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


_FMA_PEAK = None


def fma_peak(engine):
    """FP32 FMA-pipe peak measured in this process (probe library), once."""
    global _FMA_PEAK
    if _FMA_PEAK is None:
        scalar = max(engine.fma_peak_tflops(0, 1 << 13), engine.fma_peak_tflops(2, 1 << 13))
        packed = max(engine.fma_peak_tflops(1, 1 << 13), engine.fma_peak_tflops(3, 1 << 13))
        _FMA_PEAK = (max(scalar, packed), scalar, packed)
    return _FMA_PEAK


def profiled_traffic(name):
    """DRAM bytes per launch of a workload's dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum, with the profile file it came from) -- a
    number taken under the profiler, so it is quoted from the capture, never re-measured in a timed run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            rec = json.load(fh).get(name)
        return (rec["bytes_per_launch"], rec["source"]) if rec else (None, None)
    except Exception:
        return None, None


def profiled_fma_pipe(name):
    """FMA-pipe utilisation of a workload's dominant kernel from the same committed capture (sm__pipe_fma_cycles_active,
    % of active / of elapsed cycles) -- the hardware-side view of the executed-flop fraction; quoted, never re-measured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            rec = json.load(fh).get(name) or {}
        return rec.get("fma_pipe_active_pct"), rec.get("fma_pipe_elapsed_pct")
    except Exception:
        return None, None


def roofline_for(engine, w, spec, kern_ms, evals_per_kernel_call, sm_clk=0.0):
    """Roofline record of the dominant kernel: HBM bandwidth for the single large circuit, FP32 FMA otherwise."""
    flops_exec, sweeps = engine.work_model(spec)
    flops_canon = canonical_flops(w)
    kernel_name = engine.kernel_name(spec)
    traffic, traffic_src = profiled_traffic(w["name"]) if not w.get("overridden") else (None, None)
    if w["forward_only"] and engine.kernel_class(spec) == 2:
        state_bytes = (1 << w["n"]) * 8
        alg_bytes = sweeps * 2.0 * state_bytes * evals_per_kernel_call
        peak, src = measured_hbm_peak()
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        exec_tflops = flops_exec * evals_per_kernel_call / (kern_ms * 1e-3) / 1e12
        r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
             "kernel": kernel_name, "kernel_ms": kern_ms, "state_sweeps_per_eval": sweeps, "state_bytes": state_bytes,
             "algorithmic_bytes_per_eval": sweeps * 2.0 * state_bytes,
             "gate_by_gate_sweeps": w["n"] * (w["L"] + (1 if w["n_feat"] else 0)) + 1,
             "executed_tflops": exec_tflops, "executed_fp32_frac": exec_tflops / fma_peak(engine)[0], "peak_source": src}
    else:
        peak, peak_scalar, peak_packed = fma_peak(engine)
        achieved = flops_exec * evals_per_kernel_call / (kern_ms * 1e-3) / 1e12
        r = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
             # `achieved` counts the flop the kernel EXECUTES (exact algebra removes canonical work: product-state first
             # layer, rotated read-out, light-cone pruning); the canonical gate-by-gate count of SURVEY.md 8(d) is beside it
             "flop_basis": "executed", "kernel": kernel_name, "kernel_ms": kern_ms, "flop_per_eval_executed": flops_exec,
             "flop_per_eval_canonical": flops_canon,
             "canonical_equiv_tflops": flops_canon * evals_per_kernel_call / (kern_ms * 1e-3) / 1e12,
             "peak_source": "FP32 FMA microbenchmark measured in this process (oqrl_fma_peak, probe library: scalar FFMA %.1f, packed FFMA2 %.1f TFLOP/s); "
                            "MEASURED_PEAKS.json has no FP32 figure. Nominal 148 SM x 128 lanes x 2 x f_clk = %.1f TFLOP/s at the sampled %.0f MHz"
                            % (peak_scalar, peak_packed, 148 * 128 * 2 * sm_clk * 1e6 / 1e12, sm_clk)}
    if traffic_src:
        r["traffic_source"] = traffic_src
        act, ela = profiled_fma_pipe(w["name"])
        if act is not None and r["bound"] == "fp32":
            r["fma_pipe_active_pct_ncu"], r["fma_pipe_elapsed_pct_ncu"] = act, ela
    return r


def timed(torch, flush, fn, steps, warmup):
    """Median device time (ms) of fn() over `steps` launches, L2 flushed before each, CUDA events on the current stream."""
    for _ in range(max(warmup, 3)):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return statistics.median(ms), sum(ms) / len(ms)


def parity_sample(torch, engine, w, spec, params_np, cols, idx_np, fit_gpu=None, n_cand=8):
    """Parity bit of a configuration inside the bench: the CUDA result of a sample of the candidates against the C
    oracle on the same inputs (the checker role of oracle/, as in smoke())."""
    from oracle import c_oracle
    ospec = oracle_spec(w)
    if w["forward_only"]:
        # one large circuit: the oracle at the benchmarked depth takes ~13 s of one thread, so the bench compares at
        # 2 layers (same kernels, same planner) and checks the benchmarked depth through size-independent properties;
        # the full-depth oracle comparison is tests/test_gpu_parity.py::test_c5_benchmarked_spec_24_qubits_8_layers
        base, q_ref, Ls = cpu_forward_baseline(w, params_np, cols, idx_np, want_q=True)
        import dataclasses
        spec_s = dataclasses.replace(spec, n_layers=Ls)
        x = cols["obs"][idx_np[:1]]
        q = engine.qvalues(spec_s, params_np[:1, :Ls * w["n"] * 3], x).cpu().numpy()
        err = rel_err(q, q_ref)
        sv = engine.statevector(spec, params_np[0], x[0])
        prob = sv.real.double() ** 2 + sv.imag.double() ** 2
        norm = float(prob.sum())
        k = torch.arange(1 << w["n"], device=sv.device)
        qfull = engine.qvalues(spec, params_np[:1], x).cpu().numpy()[0, 0]
        zerr = max(abs(float((prob * (1 - 2 * ((k >> (w["n"] - 1 - i)) & 1)).double()).sum()) - float(qfull[i])) for i in range(w["n_out"]))
        ok = err < TOL and abs(norm - 1.0) < 1e-4 and zerr < 1e-4
        return {"ok": bool(ok), "oracle_rel_err": err, "oracle_layers": Ls, "unitarity_err": abs(norm - 1.0), "readout_vs_state_err": zerr,
                "checked": f"Q-values vs C oracle at {Ls} of {w['L']} layers; norm and <Z_i> of the stored state at {w['L']} layers"}, base
    pick = np.random.default_rng(7).choice(w["P"], size=min(n_cand, w["P"]), replace=False)
    obs, nxt, act, rew, term = (cols[k][idx_np] for k in ("obs", "next_obs", "act", "rew", "term"))
    ref = c_oracle.fitness(ospec, params_np[pick], obs, act, rew, nxt, term)
    got = fit_gpu[pick]
    err = rel_err(got, ref)
    same_rank = bool(np.array_equal(np.argsort(got)[::-1], np.argsort(ref)[::-1]))
    return {"ok": bool(err < TOL and same_rank), "oracle_rel_err": err, "ranking_identical": same_rank,
            "checked": f"fitness of {len(pick)} sampled candidates x {w['T']} transitions vs C oracle"}, None


def dataset_upload_timing(torch, engine, w, cols, n_rows=100000):
    """One-time residency of a table the size of the real dataset (offline_cartpole_test_v5.hdf5 has 100 000 rows; the
    file itself is not on the GPU box, so the committed 4096 rows are tiled): host arrays -> oqrl_dataset_upload ->
    HBM, including the half-angle table kernel.  Wall clock, median of 3."""
    reps = -(-n_rows // cols["obs"].shape[0])
    big = {k: np.ascontiguousarray(np.concatenate([v] * reps)[:n_rows]) for k, v in cols.items()}
    times, ds = [], None
    for _ in range(3):
        del ds
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ds = engine.DeviceDataset(big["obs"], big["next_obs"], big["act"], big["rew"], big["term"])
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    nf = w["n_feat"]
    out = {"rows": n_rows, "ms": 1e3 * statistics.median(times), "h2d_bytes": int(ds.h2d_bytes),
           "resident_bytes": int(n_rows * (2 * nf * 4 + nf * 16 + 16)),
           "what": "observations + next observations (f32), half-angle table cos/sin(x/2) of both (float4 per feature), {reward, terminal, action} (float4)"}
    del ds
    return out


def also_workload(torch, engine, args, name, dev, flush):
    """Short run of another BASELINE configuration for the `also` map (N = 1 only)."""
    from oqrl_b200 import CircuitSpec
    w = workload(args, name)
    spec = CircuitSpec(w["n"], w["L"], w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])
    params_np, cols, idx_np = synthetic_inputs(w, 0)
    ds = engine.DeviceDataset(cols["obs"], cols["next_obs"], cols["act"], cols["rew"], cols["term"])
    d_params = torch.tensor(params_np, device=dev)
    d_idx = torch.tensor(idx_np, device=dev)
    d_x = torch.tensor(cols["obs"][idx_np], device=dev)
    fwd = w["forward_only"]
    call = (lambda: engine.qvalues(spec, d_params, d_x)) if fwd else (lambda: engine.population_fitness(spec, d_params, ds, d_idx))
    evals = (1.0 if fwd else 2.0) * w["P"] * w["T"]
    steps = {"c1": 200, "c3": 3, "c4": 3, "c5": 30}.get(name, 10)
    med, mean = timed(torch, flush, call, steps, 3)
    out = {"workload": w["desc"], "value": evals / (mean * 1e-3), "unit": UNIT, "ms_per_step": mean, "steps": steps,
           "roofline": roofline_for(engine, w, spec, med, evals)}
    fit = None if fwd else call().cpu().numpy()
    out["parity"], base = parity_sample(torch, engine, w, spec, params_np, cols, idx_np, fit, n_cand=8 if name != "c4" else 2)
    if not args.no_cpu_baseline:
        if fwd:
            out["cpu_baseline"] = base
        else:
            rate, cores, sample_desc = cpu_port_rate(w, params_np, cols, idx_np, 3.0)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc + "; C oracle (fp64, OpenMP over candidates)"}
    del ds
    return out


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
    w["overridden"] = bool(args.population or args.transitions or args.layers or args.qubits or args.entangler or args.general_kernel)
    params_np, cols, idx_np = synthetic_inputs(w, rank)
    P, T = w["P"], w["T"]
    fwd = w["forward_only"]

    # one-time residency: the whole transition table goes to HBM (dataset.py -> oqrl_dataset_upload)
    ds = engine.DeviceDataset(cols["obs"], cols["next_obs"], cols["act"], cols["rew"], cols["term"])
    d_params = torch.tensor(params_np, device=dev)
    d_idx = torch.tensor(idx_np, device=dev)
    d_x = torch.tensor(cols["obs"][idx_np], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    collective = "none"
    exchange_verified = None
    sharded_ga_ok = None
    p2p = None
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
        if world > 1:
            # correctness bit of the exchange: every rank's parameter block goes to rank 0 (NCCL, outside any timed
            # region), which evaluates the WHOLE population on its own GPU and compares the exchanged vector bit for bit
            gathered = step().clone()
            all_params = [torch.empty_like(d_params) for _ in range(world)] if rank == 0 else None
            dist.gather(d_params, all_params, dst=0)
            ok = torch.ones(1, device=dev)
            if rank == 0:
                whole = engine.population_fitness(spec, torch.cat(all_params), ds, d_idx)
                ok[0] = float(torch.equal(whole, gathered))
            dist.broadcast(ok, src=0)
            exchange_verified = bool(ok.item())
            sharded_ga_ok = sharded_ga_check(torch, dist, dev)

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
    kern_ms, _ = timed(torch, flush, kernel_call, args.steps, 0)

    # ---- the same step with selection on the device (optim.py:207: the five best of the exchanged vector): at N > 1
    # the selection kernel itself waits for the peers' flags (oqrl_ga_rank_peers), so the barrier is no launch of its own
    with_selection = None
    if not fwd:
        def sel_step():
            if p2p is not None:
                return engine.ga_rank(p2p.scatter(d_params, d_idx), 5, peers=p2p.rank_args())
            return engine.ga_rank(kernel_call() if world == 1 else evaluator(d_params), 5)
        for _ in range(5):
            sel_step()
        barrier()
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in sev:
            flush.fill_(1)
            a.record()
            sel_step()
            b.record()
        barrier()
        tot = torch.tensor([sum(a.elapsed_time(b) for a, b in sev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        sel_ms = float(tot.item()) / args.steps
        with_selection = {"ms_per_step": sel_ms, "value": evals_per_step / (sel_ms * 1e-3), "unit": UNIT,
                          "what": "fitness of the whole population + exchange + oqrl_ga_rank (k = 5) on every rank, device-timed, max over ranks"}

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
    elif p2p is not None and hasattr(p2p, "host_step"):
        # sharded host entry point (oqrl_fitness_scatter_host): pinned parameters read in place, indices staged, the
        # exchanged vector read back only by the rank that asks for it (rank 0 = the rank that runs selection)
        h2d, d2h = h_params.numel() * 4 + h_idx.numel() * 4, (P * world * 4 if rank == 0 else 0)
        e2e_path = ("oqrl_fitness_scatter_host per step on every rank: pinned host parameters read in place by the fitness kernel, row indices "
                    "staged H2D, peer stores + oqrl_peer_barrier, rank 0 copies the whole exchanged vector into pinned host memory")

        def e2e_step():
            p2p.host_step(h_params, h_idx, h_out if rank == 0 else None)
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

    # ---- config 3 with its population of 16384 sharded over the ranks (strong scaling; BASELINE.json configs[2]) ----
    c3_strong = None
    if world > 1 and not fwd and not args.no_also:
        c3_strong = c3_strong_run(torch, dist, engine, args, rank, world, dev, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    evals_per_kernel_call = evals_per_step / world
    sm_clk = (clocks or {}).get("sm_mhz") or 0.0
    roofline = roofline_for(engine, w, spec, kern_ms, evals_per_kernel_call, sm_clk)

    if engine.kernel_class(spec) == 0 and not fwd:
        # same launch with the last-layer light-cone pruning switched off: every Rot gate of every layer executed
        import dataclasses
        spec_np = dataclasses.replace(spec, no_prune=1)
        ms2, _ = timed(torch, flush, lambda: engine.population_fitness(spec_np, d_params, ds, d_idx), args.steps, 3)
        f2, _ = engine.work_model(spec_np)
        ach2 = f2 * evals_per_kernel_call / (ms2 * 1e-3) / 1e12
        roofline["unpruned_variant"] = {"kernel_ms": ms2, "flop_per_eval_executed": f2, "achieved": ach2, "frac": ach2 / roofline["peak"],
                                        "circuit_evals_per_s": evals_per_kernel_call / (ms2 * 1e-3),
                                        "note": "all gates of every layer applied (only |0>+embedding+layer 0 folded into the product state)"}

    line = {"metric": METRIC if not fwd else METRIC_FWD, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": DATA,
            "config": config_dict(w, world), "collective": collective, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "path": e2e_path},
            "gpu_launches": int(launches), "roofline": roofline}
    if with_selection is not None:
        line["with_selection"] = with_selection
    if exchange_verified is not None:
        line["exchange_verified"] = exchange_verified
    if sharded_ga_ok is not None:
        line["sharded_ga_matches_reference"] = sharded_ga_ok
    if c3_strong is not None:
        line["c3_strong"] = c3_strong
    if not args.no_cpu_baseline and world == 1:
        if fwd:
            line["cpu_baseline"] = cpu_forward_baseline(w, params_np, cols, idx_np)
        else:
            rate, cores, sample_desc = cpu_port_rate(w, params_np, cols, idx_np, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample_desc + "; C oracle (fp64, OpenMP over candidates)"}
            ref_files = cpu_reference_files_rate(w, params_np, cols, idx_np, 4.0)
            if ref_files:
                line["cpu_baseline"]["reference_files"] = ref_files
            if w["n"] <= 8:
                loop_rate = cpu_reference_loop_rate(w, params_np, cols, idx_np, 3.0)
                line["cpu_baseline"]["reference_structured_python_loop"] = {
                    "value": loop_rate, "unit": UNIT, "cores": 1,
                    "note": "per-candidate x per-sample loop (optim.py:204, agent.py:51) without PennyLane tape overhead"}
    if world == 1 and not fwd:
        line["dataset_upload"] = dataset_upload_timing(torch, engine, w, cols)
    if world == 1 and not fwd and engine.kernel_class(spec) < 2:
        # the whole GA loop on the device: oqrl_ga_run = fitness, selection, breeding per
        # generation, every generation enqueued by one library call; device-timed around the call
        G = 10
        pop = d_params.clone()
        for _ in range(3):
            engine.ga_run(spec, pop, ds, d_idx, G, 5, 5, seed=1, generation0=0)
        ga_ms, _ = timed(torch, flush, lambda: engine.ga_run(spec, pop, ds, d_idx, G, 5, 5, seed=1, generation0=0), max(3, min(args.steps, 50)), 0)
        line["ga_loop"] = {"ms_per_generation": ga_ms / G, "generations_per_call": G, "value": evals_per_step * G / (ga_ms * 1e-3), "unit": UNIT,
                           "what": "oqrl_ga_run: fused fitness kernel + oqrl_ga_rank (k = 5) + oqrl_ga_breed_random per generation, "
                                   "10 generations per library call, L2 flushed before each call"}
    if world == 1 and not args.no_also and not w["overridden"]:
        # the other configurations of BASELINE.json, short runs (seconds each), so that every one is driver-visible
        line["also"] = {}
        for name in ("c1", "c2", "c3", "c4", "c5"):
            if name != w["name"]:
                line["also"][name] = also_workload(torch, engine, args, name, dev, flush)
        if w["name"] == "c2":
            fit = kernel_call().cpu().numpy()
            line["parity"], _ = parity_sample(torch, engine, w, spec, params_np, cols, idx_np, fit, n_cand=64)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sharded_ga_check(torch, dist, dev):
    """ShardedDeviceGAOptimizer (fitness sharded over the ranks, oqrl_ga_rank_peers, breeding on every rank) must
    reproduce the GA trajectory recorded from the reference's own optim.py (tests/golden/ga_trajectory_ref.npz) bit for
    bit on every rank.  Collective; outside every timed region."""
    from oqrl_b200 import GAOptimizer, ShardedDeviceGAOptimizer, VQC
    sample = dict(np.load(os.path.join(ROOT, "tests", "golden", "cartpole_v5_sample.npz")))
    ref = dict(np.load(os.path.join(ROOT, "tests", "golden", "ga_trajectory_ref.npz")))

    class Exp:
        def __init__(self, i):
            self.obs, self.action = torch.tensor(sample["sample_obs"][i]), torch.tensor(int(sample["sample_act"][i]))
            self.reward, self.next_obs = torch.tensor(float(sample["sample_rew"][i])), torch.tensor(sample["sample_next_obs"][i])
            self.terminated = torch.tensor(float(sample["sample_term"][i]))
    ok = True
    try:
        for seed in (0, 1):
            rng = np.random.default_rng(seed)
            policy = VQC(4, 2, n_layers=2, rng=rng)
            VQC(4, 2, n_layers=2, rng=rng)
            GAOptimizer(model=policy, rng=rng)
            ga = ShardedDeviceGAOptimizer(policy, rng=rng)
            batch = [Exp(i) for i in rng.choice(4096, 16, replace=False)]
            ga.optimize(None, batch)
            pop = np.stack([i[0].cpu().numpy() for i in ga.population])
            ok &= bool(np.array_equal(pop, ref[f"seed{seed}_final_population"])) and not ga._exchange.timed_out()
    except Exception as e:          # symmetric memory unavailable etc.: report, do not kill the benchmark
        print(f"sharded GA check failed to run: {type(e).__name__}: {e}", file=sys.stderr)
        ok = False
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def c3_strong_run(torch, dist, engine, args, rank, world, dev, flush):
    """BASELINE.json configs[2]: the 8-qubit 8-layer re-uploading agent, population 16384 SHARDED over the ranks
    (P / N candidates per rank, T = 1024), fitness exchanged by the fused peer stores.  Collective on every rank."""
    from oqrl_b200 import CircuitSpec
    from oqrl_b200.sharding import PeerScatterFitness
    w = workload(args, "c3")
    P_total = w["P"]
    if P_total % world:
        return {"skipped": f"population {P_total} does not divide over {world} ranks"}
    w = dict(w, P=P_total // world)
    spec = CircuitSpec(w["n"], w["L"], w["n_out"], w["n_feat"], w["ent"], w["reup"], w["fmap"])
    params_np, cols, idx_np = synthetic_inputs(w, rank)
    ds = engine.DeviceDataset(cols["obs"], cols["next_obs"], cols["act"], cols["rew"], cols["term"])
    d_params = torch.tensor(params_np, device=dev)
    d_idx = torch.tensor(idx_np, device=dev)
    try:
        ex = PeerScatterFitness(w["P"], spec, ds)
    except Exception as e:
        return {"skipped": f"peer exchange unavailable: {type(e).__name__}"}

    def step():
        return ex(d_params, d_idx)
    steps = 5
    for _ in range(2):
        step()
    dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        step()
        b.record()
    dist.barrier(); torch.cuda.synchronize()
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item()) / steps
    evals = 2.0 * P_total * w["T"]
    # end to end: pinned host parameters / indices in, exchanged vector out on rank 0
    h_params = torch.tensor(params_np).pin_memory()
    h_idx = torch.tensor(idx_np).pin_memory()
    h_out = torch.empty(P_total, dtype=torch.float32).pin_memory()
    if hasattr(ex, "host_step"):
        def e2e_step():
            ex.host_step(h_params, h_idx, h_out if rank == 0 else None)
    else:
        def e2e_step():
            h_out.copy_(ex(h_params.to(dev, non_blocking=True), h_idx.to(dev, non_blocking=True)), non_blocking=True)
            torch.cuda.synchronize()
    e2e_step()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    dist.barrier(); torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    # correctness of the sharded result: a sample of this rank's candidates against the C oracle
    out = {"workload": WORKLOADS["c3"][9], "scaling": "strong", "population_total": P_total, "population_per_gpu": w["P"], "transitions": w["T"],
           "value": evals / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "e2e": {"value": evals * steps / float(el.item()), "unit": UNIT}}
    if rank == 0:
        fit = step().cpu().numpy()[: w["P"]]
        out["parity"], _ = parity_sample(torch, engine, w, spec, params_np, cols, idx_np, fit, n_cand=8)
    del ex, ds
    return out


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
