/*
 * oqrl_b200_probe -- measurement and test hooks of the B200-native PQC fitness path.
 *
 * A SEPARATE library (liboqrl_b200_probe.so, built from oqrl_b200/csrc/probe.cu; links liboqrl_b200.so).  Nothing
 * here replaces a reference interface and nothing on the product path calls it: bench.py uses the microbenchmarks
 * as roofline denominators and the work models for executed-flop / sweep counts, tests/hbm_plan_emulator.py reads
 * the HBM-class pass list.  Same conventions as oqrl_b200.h (0 / negative status, oqrl_last_error()).
 */
#ifndef OQRL_B200_PROBE_H
#define OQRL_B200_PROBE_H

#include "oqrl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Executed (not canonical) FP32 flop per circuit evaluation of the kernel a spec
 * dispatches to, and HBM state sweeps per circuit for the HBM-resident class. */
int oqrl_work_model(const oqrl_spec *spec, double *flops_per_eval, double *state_sweeps);
/* Test hook: the host-built pass list of the HBM-resident class for `spec` (binary array of the
 * library's internal pass records; layout mirrored in tests/hbm_plan_emulator.py).  `buf` may be
 * NULL to query the size.  keep_state != 0 builds the plan oqrl_statevector uses. */
int oqrl_hbm_plan(const oqrl_spec *spec, void *buf, uint64_t capacity, uint64_t *n_bytes, int32_t keep_state);
/* Test hook: the work-unit list the 1..4-qubit fitness kernel deals to `n_warps` persistent warps for a population of
 * n_cand candidates on n_trans transitions (host logic of pqc_reg.cu, checked on the CPU by tests/test_reg_plan.py):
 * out7 = {n_full, n_big, big_per_cand, seg_big, seg_rest, n_units, n_segs} -- units [0, n_full) are whole candidates,
 * the next n_big units cover seg_big 256-transition segments each (big_per_cand per remaining candidate), then one unit
 * of seg_rest segments per remaining candidate.  can_split = 0: the launch has no partial-sum scratch. */
int oqrl_reg_plan(int32_t n_cand, int32_t n_trans, int32_t n_warps, int32_t can_split, int64_t partial_capacity, int32_t *out7);
/* FP32 FMA-pipe peak microbenchmark.  variant bit 0: 0 = scalar FFMA, 1 = packed FFMA2;
 * variant >> 1 = operand pattern (0: one fresh register operand per instruction, 1: two,
 * 2: three, 3: two plus a warp-uniform multiplier).  Runs `iters` rounds of 16 independent
 * FMA chains per thread on every SM and returns achieved TFLOP/s measured with CUDA events
 * on `stream`.  Variants 0/1 are the roofline denominator.  Variants 32..39 are the instruction-
 * fetch probe: an unrolled loop body of 256/384/512/768/1024/1536/2048/3072 packed FMAs
 * (16 bytes each) run by staggered warps; bits 8..15 select resident warps per sub-partition. */
int oqrl_fma_peak(int32_t variant, int32_t iters, double *tflops, void *stream);

/* Arithmetic rate of one register block of the HBM-resident class on registers alone (16 amplitudes, four SU(2)
 * gates per iteration, no memory traffic), in TFLOP/s at 28 flop per amplitude pair and gate.  form 0 = the
 * complex-packed body of pqc_hbm.cu, 1 = scalar FFMA, 2 = two-circuit packing (the small-n families).  Run with
 * `warps_per_smsp` resident warps per SM sub-partition. */
int oqrl_gate_rate(int32_t form, int32_t warps_per_smsp, int32_t iters, double *tflops, void *stream);

/* Exposed cost of re-tiling one 16-qubit state (512 KiB = eight 64 KiB tiles, every tile's 256-byte rows spread evenly
 * over the eight tiles of the next pass) in the media a B200 offers: medium 0 = distributed shared memory of an
 * eight-CTA cluster with bulk copies, 1 = L2 / HBM with bulk stores + bulk loads (what the pass kernel does), 2 =
 * distributed shared memory with st.shared::cluster stores.  18 clusters run `iters` exchanges back to back; returns
 * microseconds per exchange and the aggregate payload rate.  The measurement behind DESIGN.md's choice for config 4. */
int oqrl_exchange_rate(int32_t medium, int32_t iters, double *us_per_exchange, double *gbytes_per_s, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OQRL_B200_PROBE_H */
