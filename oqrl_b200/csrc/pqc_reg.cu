// Register-resident PQC kernels for n <= 4 qubits (BASELINE configs 1 and 2).
//
// Replaces the per-sample QNode loop of /root/reference/src/agent.py:48-61 and
// the per-candidate loop body of /root/reference/src/optim.py:64-173.
//
// Work decomposition (one thread = one transition, both circuits):
//   * a CTA belongs to one candidate, so every gate coefficient is CTA-uniform
//     and lives in shared memory as four scalars per gate (one broadcast LDS.128), used as the
//     scalar-broadcast operand of FFMA2;
//   * a thread owns the full 2^n-amplitude state of BOTH circuits of a
//     transition -- Q(s) in the low half and Q(s') in the high half of packed
//     fp32x2 registers -- so every gate is 16 FFMA2/FMUL2 per amplitude pair
//     for two circuits, with zero shuffles and zero shared-memory state traffic;
//   * the CNOT ring is a compile-time renaming of registers (0 instructions)
//     when the layer count is a template constant;
//   * <Z_i>, the TD error and its square are computed in registers; the sum
//     over transitions is a fixed-order shuffle tree in fp64 (deterministic).
// State preparation folds |0..0> and the first AngleEmbedding into a product
// state (cheaper than the canonical 6*2^n flop per RY; accounted as executed
// flop in reg_flops_per_eval, never as canonical work).
#include "common.cuh"

namespace oqrl {
namespace {

constexpr int kRegThreads = 128;   // forward-only kernel CTA size
constexpr int kRegMaxGates = 256;  // L * n bound for this family
constexpr int kRegSegLen = 256;    // transitions per reduction segment (8 lane-iterations)
constexpr int kRegStageSmemMax = 225 * 1024;   // shared memory a staged fitness launch may take (tables + batch)

template <int N>
struct RegState {
    u64 re[1 << N];
    u64 im[1 << N];
};

template <int N>
__device__ __forceinline__ void apply_su2(RegState<N> &st, const Su2Scalar &g, int wire_const)
{
    // wire_const is a compile-time constant after unrolling.
    const int pos = N - 1 - wire_const;
#pragma unroll
    for (int k = 0; k < (1 << N); ++k) {
        if (((k >> pos) & 1) == 0) {
            const int k1 = k | (1 << pos);
            su2_pair(g, st.re[k], st.im[k], st.re[k1], st.im[k1]);
        }
    }
}

template <int N>
__device__ __forceinline__ void apply_ry(RegState<N> &st, u64 c, u64 s, int wire_const)
{
    const int pos = N - 1 - wire_const;
    const u64 ns = fneg2(s);
#pragma unroll
    for (int k = 0; k < (1 << N); ++k) {
        if (((k >> pos) & 1) == 0) {
            const int k1 = k | (1 << pos);
            ry_pair(c, s, ns, st.re[k], st.im[k], st.re[k1], st.im[k1]);
        }
    }
}

template <int N, int R>
__device__ __forceinline__ void permute_sel(RegState<N> &st)
{
    RegState<N> t;
#pragma unroll
    for (int k = 0; k < (1 << N); ++k) {
        t.re[sel_perm(k, N, R)] = st.re[k];
        t.im[sel_perm(k, N, R)] = st.im[k];
    }
    st = t;
}

template <int N>
__device__ __forceinline__ void entangle_sel_runtime(RegState<N> &st, int r)
{
    if constexpr (N >= 2) {
        if (N == 2 || r == 1) permute_sel<N, 1>(st);
        if constexpr (N >= 3) { if (r == 2) permute_sel<N, 2>(st); }
        if constexpr (N >= 4) { if (r == 3) permute_sel<N, 3>(st); }
    }
}

template <int N>
__device__ __forceinline__ void entangle_cz(RegState<N> &st)
{
#pragma unroll
    for (int k = 0; k < (1 << N); ++k) {
        if (cz_ring_negative(k, N)) {
            st.re[k] = fneg2(st.re[k]);
            st.im[k] = fneg2(st.im[k]);
        }
    }
}

// State preparation.  |0..0>, the AngleEmbedding and -- when the circuit has at least one layer --
// the first layer's Rot gates act wire by wire on a product state, so the register state is built as
// the tensor product of per-wire complex 2-vectors  v_w = Rot_w (cos(x_w/2), sin(x_w/2))^T :
// 8 packed ops per wire + one complex multiply per amplitude per doubling step, instead of
// n * 2^(n-1) general pair updates.  (Executed flop are accounted in reg_flops_per_eval; the
// canonical gate-by-gate count is reported separately by bench.py.)
// L0: 1 = layer 0 present (compile time), 0 = absent, -1 = decided at run time by `layer0 != nullptr`
template <int N, int L0>
__device__ __forceinline__ void prepare_product(RegState<N> &st, const Su2Scalar *__restrict__ layer0_rt,
                                                const u64 (&cw)[N], const u64 (&sw)[N])
{
    const bool layer0 = L0 > 0 ? true : (L0 == 0 ? false : layer0_rt != nullptr);
    const Su2Scalar *__restrict__ layer0_tab = layer0_rt;
    u64 v0r[N], v0i[N], v1r[N], v1i[N], nv0i[N], nv1i[N];
#pragma unroll
    for (int w = 0; w < N; ++w) {
        if (layer0) {
            const Su2Scalar g = layer0_tab[w];
            // [[a, b], [-conj(b), conj(a)]] (c, s)^T
            v0r[w] = ffma2(bcast2(g.br), sw[w], fmul2(bcast2(g.ar), cw[w]));
            v0i[w] = ffma2(bcast2(g.bi), sw[w], fmul2(bcast2(g.ai), cw[w]));
            v1r[w] = ffma2(bcast2(g.ar), sw[w], fmul2(bcast2(-g.br), cw[w]));
            v1i[w] = ffma2(bcast2(-g.ai), sw[w], fmul2(bcast2(g.bi), cw[w]));
            nv0i[w] = fneg2(v0i[w]);     // sign flips run on the integer pipe, not the FMA pipe
            nv1i[w] = fneg2(v1i[w]);
        } else {
            v0r[w] = cw[w]; v0i[w] = 0ull; v1r[w] = sw[w]; v1i[w] = 0ull; nv0i[w] = 0ull; nv1i[w] = 0ull;
        }
    }
    if constexpr (N == 4 && L0 > 0) {
        // (v_0 x v_1) x (v_2 x v_3): 4 + 4 + 16 complex multiplies instead of 4 + 8 + 16
        u64 ar[4], ai[4], br[4], bi[4], nbi[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const u64 xr = j ? v1r[0] : v0r[0], xi = j ? v1i[0] : v0i[0];
            const u64 yr = j ? v1r[2] : v0r[2], yi = j ? v1i[2] : v0i[2];
            ar[2 * j] = ffma2(xi, nv0i[1], fmul2(xr, v0r[1]));     ai[2 * j] = ffma2(xi, v0r[1], fmul2(xr, v0i[1]));
            ar[2 * j + 1] = ffma2(xi, nv1i[1], fmul2(xr, v1r[1])); ai[2 * j + 1] = ffma2(xi, v1r[1], fmul2(xr, v1i[1]));
            br[2 * j] = ffma2(yi, nv0i[3], fmul2(yr, v0r[3]));     bi[2 * j] = ffma2(yi, v0r[3], fmul2(yr, v0i[3]));
            br[2 * j + 1] = ffma2(yi, nv1i[3], fmul2(yr, v1r[3])); bi[2 * j + 1] = ffma2(yi, v1r[3], fmul2(yr, v1i[3]));
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) nbi[b] = fneg2(bi[b]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                st.re[4 * a + b] = ffma2(ai[a], nbi[b], fmul2(ar[a], br[b]));
                st.im[4 * a + b] = ffma2(ai[a], br[b], fmul2(ar[a], bi[b]));
            }
        }
        return;
    }
    st.re[0] = v0r[0]; st.im[0] = v0i[0];
    if constexpr (N >= 1) { st.re[1] = v1r[0]; st.im[1] = v1i[0]; }
    // grow one wire at a time: wire w becomes the new least significant bit
#pragma unroll
    for (int w = 1; w < N; ++w) {
#pragma unroll
        for (int k = (1 << w) - 1; k >= 0; --k) {
            const u64 pr = st.re[k], pi = st.im[k];
            if (layer0) {
                st.re[2 * k] = ffma2(pi, nv0i[w], fmul2(pr, v0r[w]));
                st.im[2 * k] = ffma2(pi, v0r[w], fmul2(pr, v0i[w]));
                st.re[2 * k + 1] = ffma2(pi, nv1i[w], fmul2(pr, v1r[w]));
                st.im[2 * k + 1] = ffma2(pi, v1r[w], fmul2(pr, v1i[w]));
            } else {
                st.re[2 * k] = fmul2(pr, v0r[w]);
                st.im[2 * k] = 0ull;
                st.re[2 * k + 1] = fmul2(pr, v1r[w]);
                st.im[2 * k + 1] = 0ull;
            }
        }
    }
}

// <Z_i> for every wire, packed (circuit lo, circuit hi).
template <int N>
__device__ __forceinline__ void readout(const RegState<N> &st, u64 (&z)[N], int n_out)
{
    u64 p[1 << N];
#pragma unroll
    for (int k = 0; k < (1 << N); ++k) p[k] = ffma2(st.im[k], st.im[k], fmul2(st.re[k], st.re[k]));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        z[i] = 0ull;
        if (i >= n_out) continue;          // warp-uniform
        const int pos = N - 1 - i;
        u64 s0 = 0ull, s1 = 0ull;
        bool f0 = true, f1 = true;
#pragma unroll
        for (int k = 0; k < (1 << N); ++k) {
            if ((k >> pos) & 1) { s1 = f1 ? p[k] : fadd2(s1, p[k]); f1 = false; }
            else                { s0 = f0 ? p[k] : fadd2(s0, p[k]); f0 = false; }
        }
        z[i] = fsub2(s0, s1);
    }
}

// Read-out with the last layer folded into the observable (readout_is_single_wire): for output i the pre-entangler
// Z-string is one wire w_i, and  <Z> after its gate U = [[a, b], [-conj b, conj a]]  is
//     n_z (S0 - S1) + G_r C_r + G_i C_i,   S0/S1 = sum |psi|^2 over the wire's 0/1 half,  C = sum conj(psi_0) psi_1,
//     n_z = |a|^2 - |b|^2,  G_r = 4 Re(conj(a) b),  G_i = -4 Im(conj(a) b)            (obs[i] = {n_z, G_r, G_i, -}).
// The |psi|^2 sums are shared between the outputs: one accumulate-in-place chain per SECTOR (the 2^NOUT value
// combinations of the read-out wires' index bits), 2 * 2^N packed FMAs for all outputs together, and S0 - S1 of an
// output is a signed sum of sectors folded into its final combine.  C_r, C_i stay per output (2 * 2^N FMAs each).
template <int N, int LC, int ENT, int NOUT>
__device__ __forceinline__ void readout_rotated(const RegState<N> &st, const Su2Scalar *__restrict__ obs, u64 (&z)[N])
{
    constexpr unsigned kAll = (1u << N) - 1u;
    int pos[NOUT];
#pragma unroll
    for (int i = 0; i < NOUT; ++i) {
        const unsigned m = readout_index_mask(N, LC, i, ENT) & kAll;
        pos[i] = 0;
#pragma unroll
        for (int b = 0; b < N; ++b)
            if ((m >> b) & 1u) pos[i] = b;
    }
    u64 sec[1 << NOUT];
    bool sec_first[1 << NOUT];
#pragma unroll
    for (int s = 0; s < (1 << NOUT); ++s) { sec[s] = 0ull; sec_first[s] = true; }
#pragma unroll
    for (int k = 0; k < (1 << N); ++k) {
        int s = 0;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) s |= ((k >> pos[i]) & 1) << i;
        sec[s] = sec_first[s] ? fmul2(st.re[k], st.re[k]) : ffma2(st.re[k], st.re[k], sec[s]);
        sec_first[s] = false;
        sec[s] = ffma2(st.im[k], st.im[k], sec[s]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        z[i] = 0ull;
        if (i >= NOUT) continue;
        u64 cr = 0ull, ca = 0ull, cb = 0ull;
        bool first = true;
#pragma unroll
        for (int k = 0; k < (1 << N); ++k) {
            if ((k >> pos[i < NOUT ? i : 0]) & 1) continue;
            const int k1 = k | (1 << pos[i < NOUT ? i : 0]);
            if (first) {
                cr = fmul2(st.re[k], st.re[k1]);
                ca = fmul2(st.re[k], st.im[k1]);
                cb = fmul2(st.im[k], st.re[k1]);
                first = false;
            } else {
                cr = ffma2(st.re[k], st.re[k1], cr);
                ca = ffma2(st.re[k], st.im[k1], ca);
                cb = ffma2(st.im[k], st.re[k1], cb);
            }
            cr = ffma2(st.im[k], st.im[k1], cr);
        }
        const Su2Scalar o = obs[i];
        u64 q = fmul2(bcast2(o.ar), sec[0]);                 // sector 0 has every read-out bit clear: sign +
#pragma unroll
        for (int s = 1; s < (1 << NOUT); ++s)
            q = ffma2(bcast2(((s >> (i < NOUT ? i : 0)) & 1) ? -o.ar : o.ar), sec[s], q);
        q = ffma2(bcast2(o.ai), cr, q);
        q = ffma2(bcast2(o.br), ca, q);
        q = ffma2(bcast2(-o.br), cb, q);
        z[i] = q;
    }
}

// Whole circuit for two packed inputs.  LC > 0: layer count is the template
// constant (CNOT ring = register renaming); LC == 0: runtime layer loop.
// NOUT > 0: n_out is the template constant and the last layer's light cone is resolved at compile time
// (no branches); NOUT == 0: runtime n_out / runtime wire mask; NOUT < 0: |NOUT| outputs, no pruning (measurement aid).
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `after_prepare` runs right after the product state has consumed the input trigonometry: the fitness kernel
// issues the next transition's loads there, into the same registers, so the software pipeline needs no copies.
template <int N, int LC, int ENT, int NOUT, class Hook = NoHook>
__device__ __forceinline__ void run_circuit(const Su2Scalar *__restrict__ coef, int n_layers, int reupload_rt,
                                            int n_out, unsigned last_mask,
                                            const u64 (&cw)[N], const u64 (&sw)[N], const bool (&emb)[N],
                                            u64 (&z)[N], Hook after_prepare = Hook())
{
    RegState<N> st;
    // specialised read-out: the last layer and its entangler are folded into rotated observables
    constexpr bool kRotated = NOUT > 0 && LC >= 2 && readout_is_single_wire(N, LC, NOUT > 0 ? NOUT : 1, ENT);
    const int n_out_eff = NOUT > 0 ? NOUT : (NOUT < 0 ? -NOUT : n_out);
    if constexpr (NOUT > 0) last_mask = readout_wire_mask(N, LC, NOUT, ENT);
    if constexpr (NOUT < 0) last_mask = 0xffffffffu;
    prepare_product<N, (LC > 0 ? 1 : -1)>(st, (LC > 0 || n_layers > 0) ? coef : nullptr, cw, sw);   // includes layer 0's Rot gates
    after_prepare();
    // The unrolled variants are built for the reference circuit (single embedding);
    // re-uploading specs take the runtime-layer variant.
    const bool reupload = (LC > 0) ? false : (reupload_rt != 0);
    if constexpr (LC > 0) {
#pragma unroll
        for (int l = 0; l < (kRotated ? LC - 1 : LC); ++l) {
            if (l > 0) {
                // last layer: only wires inside the read-out light cone (readout_wire_mask) need their gate
#pragma unroll
                for (int w = 0; w < N; ++w)
                    if (l < LC - 1 || ((last_mask >> w) & 1u)) apply_su2<N>(st, coef[l * N + w], w);
            }
            if constexpr (N >= 2) {
                if constexpr (ENT == OQRL_ENT_SEL_CNOT) {
                    // sel_range(N, l) is a constant after unrolling
                    const int r = sel_range(N, l);
                    if (r == 1) permute_sel<N, 1>(st);
                    if constexpr (N >= 3) { if (r == 2) permute_sel<N, 2>(st); }
                    if constexpr (N >= 4) { if (r == 3) permute_sel<N, 3>(st); }
                } else {
                    entangle_cz<N>(st);
                }
            }
        }
    } else {
        for (int l = 0; l < n_layers; ++l) {
            if (l > 0) {
                const bool last = l == n_layers - 1;
                if (reupload) {
#pragma unroll
                    for (int w = 0; w < N; ++w)
                        if (emb[w] && (!last || ((last_mask >> w) & 1u))) apply_ry<N>(st, cw[w], sw[w], w);
                }
#pragma unroll
                for (int w = 0; w < N; ++w)
                    if (!last || ((last_mask >> w) & 1u)) apply_su2<N>(st, coef[l * N + w], w);
            }
            if constexpr (N >= 2) {
                if constexpr (ENT == OQRL_ENT_SEL_CNOT) entangle_sel_runtime<N>(st, sel_range(N, l));
                else entangle_cz<N>(st);
            }
        }
    }
    if constexpr (kRotated) readout_rotated<N, LC, ENT, NOUT>(st, coef + LC * N, z);
    else readout<N>(st, z, n_out_eff);
}

// Prologue: candidate's Rot gates -> shared memory.  Three half-angle sincos per gate, spread over the lanes one
// angle each (a 2-layer agent: 24 lanes busy once, instead of 10 lanes running three sincosf in a row) and parked
// in `scr`; a gate's four coefficients are then three products of those.  Same arguments, same sincosf: the
// coefficient bits equal rot_su2's.  Callers synchronise (warp or CTA) after the call.
// `angle(j, k)` returns the k-th item of this thread, item j = 3 * gate + {0: theta, 1: phi + omega, 2: phi - omega}.
template <class Angle, class Sync>
__device__ __forceinline__ void build_angle_table(float2 *scr, int n_gates, int tid, int n_threads, Angle angle, Sync sync)
{
    int k = 0;
    for (int j = tid; j < 3 * n_gates; j += n_threads, ++k) {
        float2 v;                                           // {cos, sin}
        sincosf(0.5f * angle(j, k), &v.y, &v.x);
        scr[j] = v;
    }
    sync();
}
// the raw parameters behind item j: {theta, 0} or {phi, omega}
__device__ __forceinline__ float2 angle_operands(const float *__restrict__ params, int j)
{
    const int g = j / 3, c = j - 3 * g;
    return c == 0 ? make_float2(__ldg(params + 3 * g + 1), 0.f) : make_float2(__ldg(params + 3 * g), __ldg(params + 3 * g + 2));
}
__device__ __forceinline__ float angle_of(float2 ops, int j) { return (j % 3) == 2 ? ops.x - ops.y : ops.x + ops.y; }
__device__ __forceinline__ Su2Scalar gate_from_angles(const float2 *scr, int g)
{
    const float2 t = scr[3 * g], p = scr[3 * g + 1], q = scr[3 * g + 2];
    Su2Scalar c;
    c.ar = p.x * t.x;
    c.ai = -p.y * t.x;
    c.br = -q.x * t.y;
    c.bi = -q.y * t.y;
    return c;
}

// Gate table plus, appended to it for the specialised kernels, the rotated read-out observables (readout_rotated)
// -- from the same parked angles, so the observables cost no second round of sincos.
template <int N, int LC, int ENT, int NOUT, class Angle, class Sync>
__device__ __forceinline__ void build_tables(Su2Scalar *coef, float2 *scr, int n_gates, int tid, int n_threads, Angle angle, Sync sync)
{
    constexpr bool kRotated = NOUT > 0 && LC >= 2 && readout_is_single_wire(N, LC, NOUT > 0 ? NOUT : 1, ENT);
    build_angle_table(scr, n_gates, tid, n_threads, angle, sync);
    const int n_entries = kRotated ? LC * N + NOUT : n_gates;
    for (int e = tid; e < n_entries; e += n_threads) {
        int g = e;
        if constexpr (kRotated) {
            if (e >= LC * N) {
                g = 0;
#pragma unroll
                for (int i = 0; i < (NOUT > 0 ? NOUT : 1); ++i) {
                    const unsigned m = readout_index_mask(N, LC, i, ENT);
                    int pos = 0;
#pragma unroll
                    for (int b = 0; b < N; ++b)
                        if ((m >> b) & 1u) pos = b;
                    if (e - LC * N == i) g = (LC - 1) * N + (N - 1 - pos);
                }
            }
        }
        Su2Scalar c = gate_from_angles(scr, g);
        if (kRotated && e >= LC * N) {
            const float ar = c.ar, ai = c.ai, br = c.br, bi = c.bi;
            c.ar = ar * ar + ai * ai - br * br - bi * bi;
            c.ai = 4.f * (ar * br + ai * bi);
            c.br = -4.f * (ar * bi - ai * br);
            c.bi = 0.f;
        }
        coef[e] = c;
    }
}

// Work units of one fitness launch (planned on the host, plan_units): first the whole candidates, then -- so that the
// end of the launch is made of short units -- the remaining candidates cut into `big_per_cand` units of `seg_big`
// 256-transition segments each and one unit with the `seg_rest` segments left over.
struct RegUnits {
    int n_full;         // units [0, n_full): candidate u, every segment
    int n_big;          // units [n_full, n_full + n_big): candidate n_full + v / big_per_cand, segments (v % big_per_cand) * seg_big ...
    int big_per_cand;
    int seg_big;
    int seg_rest;       // units after that: candidate n_full + w, the last seg_rest segments (0: no such units)
    int n_units;
    int n_segs;         // segments per candidate
};

// PERSISTENT warps running a host-planned list of work units; the warps of a CTA are independent (no CTA barrier after the
// prologue) and sit one per SM sub-partition in turn, so every sub-partition carries the same number of resident
// warps (single-warp CTAs at 14 per SM left two sub-partitions with four warps and two with three, and the launch
// waited for the slow pair -- profiles/r2_sweep_pt.txt).  A unit is (candidate, range of segments): the warp builds
// the candidate's table in its own shared-memory slice and runs the transitions 32 at a time.  Units that cover part
// of a candidate leave one fp64 sum per segment in `partial`; the last one to arrive adds them in segment order
// (deterministic) and writes the fitness.
//
// STAGED (the generation's batch fits in shared memory, 80 bytes per transition): ONE CTA per SM gathers the batch
// once -- index, range check, the transition's half-angle record and TD metadata -- into a structure-of-arrays copy
// in shared memory, and every iteration of every unit reads it back with conflict-free LDS.128.  Without it each
// warp gathers the same rows again for every candidate: 32 different cache lines per LDG.128, 16 data-pipe wavefronts
// per request, the L1 data pipe at 60 % and the first FMUL2 of an iteration waiting on the long scoreboard for 13 %
// of all warp samples (profiles/r2_reg_fitness_v9_ncu.txt).  Larger batches keep the direct loads (four CTAs of four
// warps per SM).
#ifndef OQRL_REG_STAGED_WARPS
#define OQRL_REG_STAGED_WARPS 16
#endif
template <int NOUT, bool STAGED> struct RegCfg {
    // the specialised reference-agent kernels fit 128 registers (16 warps per SM); the run-time-loop kernels need ~200
    static constexpr int kWarps = STAGED ? (NOUT != 0 ? OQRL_REG_STAGED_WARPS : 8) : 4;
    static constexpr int kMinCtas = STAGED ? 1 : (NOUT != 0 ? 4 : 2);
};
template <int N, int LC, int ENT, int NOUT, bool STAGED>
__global__ void __launch_bounds__(32 * RegCfg<NOUT, STAGED>::kWarps, RegCfg<NOUT, STAGED>::kMinCtas)
reg_fitness_kernel(oqrl_spec sp, const float *__restrict__ params, TransitionView tv, float gamma,
                   RegUnits wu, FitnessOut fitness, double *__restrict__ partial,
                   unsigned *__restrict__ arrivals)
{
    constexpr int kWarps = RegCfg<NOUT, STAGED>::kWarps;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const int n_gates = sp.n_layers * N;
    // per-warp slice: (n_layers + 1) * N coefficient sets, then 3 * n_gates parked angles
    const size_t slice = ((size_t)(sp.n_layers + 1) * N * sizeof(Su2Scalar) + (size_t)3 * n_gates * sizeof(float2) + 15) & ~(size_t)15;
    Su2Scalar *coef = reinterpret_cast<Su2Scalar *>(smem_raw + warp * slice);
    float2 *scr = reinterpret_cast<float2 *>(coef + (sp.n_layers + 1) * N);
    // staged batch, in blocks of 32 transitions: [meta x 32][half-angle record of wire 0 x 32] ... [wire N-1 x 32], so a
    // lane walks it with ONE pointer and compile-time offsets; ceil(T / 32) blocks (the tail zero-filled) + one spare
    // block that the software pipeline may read and never uses
    constexpr int kBlk = (N + 1) * 32;   // float4 per block
    float4 *stage = reinterpret_cast<float4 *>(smem_raw + kWarps * slice);

    bool emb[N];
    int feat[N];
#pragma unroll
    for (int w = 0; w < N; ++w) {
        feat[w] = embed_feature(w, sp.n_feat, sp.feature_map);
        emb[w] = feat[w] >= 0;
    }
    // sp.reserved bit 0 (set by tests/bench only) disables the light-cone pruning of the last layer
    const unsigned last_mask = (sp.reserved & 1) ? 0xffffffffu : readout_wire_mask(N, sp.n_layers, sp.n_out, ENT);
    // NOUT != 0 marks the fully specialised reference circuit (|NOUT| outputs; NOUT < 0 = without pruning): every wire embedded, feature index = wire
    // (the dispatcher checks n_feat == N), so the record is N consecutive float4 and nothing is predicated.
    constexpr bool kAllEmb = NOUT != 0;

    struct Unit { int cand, seg0, nseg, n_arrive; bool live; };
    auto decode = [&](unsigned u) {
        Unit x;
        x.live = u < (unsigned)wu.n_units;
        x.n_arrive = 1;
        if ((int)u < wu.n_full || !x.live) { x.cand = x.live ? (int)u : 0; x.seg0 = 0; x.nseg = wu.n_segs; }
        else {
            const int v = (int)u - wu.n_full;
            x.n_arrive = wu.big_per_cand + (wu.seg_rest > 0 ? 1 : 0);
            if (v < wu.n_big) {
                const int c = v / wu.big_per_cand;
                x.cand = wu.n_full + c; x.seg0 = (v - c * wu.big_per_cand) * wu.seg_big; x.nseg = wu.seg_big;
            } else {
                x.cand = wu.n_full + (v - wu.n_big); x.seg0 = wu.big_per_cand * wu.seg_big; x.nseg = wu.seg_rest;
            }
        }
        return x;
    };
    // the NEXT unit's parameter lines are pulled into L2 while the current unit runs (fetching them into registers
    // instead measured slower: 81.9 vs 79.9 us on config 2, 122.9 vs 114.5 us through the host entry point)
    auto fetch = [&](int cand) {
        const float *pc = params + (size_t)cand * n_gates * 3;
        if (lane * 32 < n_gates * 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + lane * 32));
    };
    auto build = [&](int cand) {
        __syncwarp();                        // every lane is done with the previous unit's table
        const float *pc = params + (size_t)cand * n_gates * 3;
        build_tables<N, LC, ENT, NOUT>(coef, scr, n_gates, lane, 32, [&](int j, int) { return angle_of(angle_operands(pc, j), j); },
                                       [] { __syncwarp(); });
        __syncwarp();
    };

    // Units are dealt out statically: the warps of the grid are numbered CTA-minor (slot = warp * gridDim.x + CTA), and
    // slot s runs units s, s + W, s + 2W, ...  A last round that fills only part of the grid therefore still loads
    // every SM sub-partition with the same number of warps (warp w of a CTA sits on sub-partition w mod 4).
    const unsigned n_slots = gridDim.x * kWarps;
    unsigned u = (unsigned)warp * gridDim.x + blockIdx.x;
    // Prologue: the first unit's table is built while the first round of batch indices is in flight; then the CTA
    // gathers the batch (STAGED) and meets at the kernel's only CTA barrier.
    int row0 = 0;
    if constexpr (STAGED) row0 = ((int)threadIdx.x < tv.n_trans && tv.idx) ? __ldg(tv.idx + threadIdx.x) : (int)threadIdx.x;
    Unit cur = decode(u);
    if (cur.live) build(cur.cand);
    bool bad_batch = false;              // an index outside the dataset: NaN fitness for every candidate of the launch
    if constexpr (STAGED) {
        bool bad = false;
        const int n_pad = ((tv.n_trans + 31) >> 5) << 5;
        for (int t = (int)threadIdx.x; t < n_pad; t += 32 * kWarps) {
            float4 *dst = stage + (size_t)(t >> 5) * kBlk + (t & 31);
            float4 rec[N], meta = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < N; ++w) rec[w] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < tv.n_trans) {
                int row = t == (int)threadIdx.x ? row0 : (tv.idx ? __ldg(tv.idx + t) : t);
                row = checked_row(row, tv.n_rows, bad);
                const float4 *trig = tv.trig + (size_t)row * (kAllEmb ? N : sp.n_feat);
#pragma unroll
                for (int w = 0; w < N; ++w)
                    if (kAllEmb || emb[w]) rec[w] = __ldg(trig + (kAllEmb ? w : feat[w]));
                meta = __ldg(tv.meta + row);
            }
            dst[0] = meta;
#pragma unroll
            for (int w = 0; w < N; ++w) dst[(1 + w) * 32] = rec[w];
        }
        bad_batch = __syncthreads_or(bad) != 0;
    }

    while (cur.live) {
        const int cand = cur.cand;
        const bool whole = cur.nseg == wu.n_segs;
        const int t0 = cur.seg0 * kRegSegLen;
        const int t1 = min(tv.n_trans, (cur.seg0 + cur.nseg) * kRegSegLen);
        const int n_it = (t1 - t0 + 31) >> 5;
        // the next unit's parameters are requested now and used when this unit is done
        u += n_slots;
        {
            const Unit nxt = decode(u);
            if (nxt.live) fetch(nxt.cand);
        }

        // Software pipeline: the half-angle record of iteration i+1 is loaded (into the same registers) right after
        // the product state of iteration i has consumed the current one.  Direct loads also fetch the row index of
        // iteration i+2 there, so the dependent idx -> row -> record chain never stalls the warp.  The TD metadata is
        // only needed at the end of an iteration and is loaded at its top.
        int t = t0 + lane;
        const float4 *rec = stage + (size_t)(t0 >> 5) * kBlk + lane;   // STAGED: this lane's slot of the current block
        // (a loaded index is not touched until the next iteration: no select on it, the bound check is on `tt`)
        auto load_row = [&](int tt) { const int tc = min(tt, t1 - 1); return tv.idx ? __ldg(tv.idx + tc) : tc; };
        float4 trig_reg[N];
        auto load_trig = [&](int row) {
            const float4 *trig = tv.trig + (size_t)row * (kAllEmb ? N : sp.n_feat);
#pragma unroll
            for (int w = 0; w < N; ++w)
                if (kAllEmb || emb[w]) trig_reg[w] = __ldg(trig + (kAllEmb ? w : feat[w]));
        };
        auto load_staged = [&]() {
#pragma unroll
            for (int w = 0; w < N; ++w)
                if (kAllEmb || emb[w]) trig_reg[w] = rec[(1 + w) * 32];
        };
        bool bad_idx = bad_batch;            // direct loads: checked where the index is consumed
        int row_cur = 0, row_next = 0;
        if constexpr (STAGED) load_staged();
        else {
            row_cur = checked_row(load_row(t), tv.n_rows, bad_idx);   // clamped: lanes past the end read a valid row and discard the result
            row_next = load_row(t + 32);
            load_trig(row_cur);
        }
        // The transitions of a generation are cut into fixed SEGMENTS of kRegSegLen.  Every segment is reduced the same
        // way (fp32 per lane, fp64 shuffle tree) and a candidate's fitness is the fp64 sum of its segment sums in segment
        // order -- whether one unit ran all of them or several units ran a few each.  So the bits of a fitness value
        // do not depend on the population size, on the candidate's position in it, or on how it is sharded over GPUs.
        double total = 0.0;
        float acc = 0.f;
        int seg = cur.seg0;
        double *part = whole ? nullptr : partial + (size_t)(cand - wu.n_full) * wu.n_segs;
#pragma unroll 1
        for (int it = 0; it < n_it; ++it, t += 32) {         // warp-uniform trip count
            const bool valid = t < t1;
            const float4 meta = STAGED ? rec[0] : __ldg(tv.meta + row_cur);
            u64 cw[N], sw[N];
#pragma unroll
            for (int w = 0; w < N; ++w) {
                if (kAllEmb || emb[w]) {
                    cw[w] = pack2(trig_reg[w].x, trig_reg[w].y);
                    sw[w] = pack2(trig_reg[w].z, trig_reg[w].w);
                } else {
                    cw[w] = pack2(1.f, 1.f);
                    sw[w] = 0ull;
                }
            }
            u64 z[N];
            run_circuit<N, LC, ENT, NOUT>(coef, sp.n_layers, sp.reupload, sp.n_out, last_mask, cw, sw, emb, z, [&]() {
                if constexpr (STAGED) {
                    rec += kBlk;                 // the block after the last one is the spare: loaded, never used
                    load_staged();
                } else {
                    row_next = checked_row(row_next, tv.n_rows, bad_idx);   // loaded one iteration ago: no stall
                    if (t + 32 < t1) load_trig(row_next);
                    row_cur = row_next;
                    row_next = load_row(t + 64);
                }
            });

            const int act = __float_as_int(meta.z);
            float qa = 0.f, mx = -INFINITY;
            constexpr int kOut = NOUT > 0 ? NOUT : -NOUT;
            const int n_out = NOUT != 0 ? kOut : sp.n_out;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (NOUT != 0 && i >= kOut) continue;
                float lo, hi;
                unpack2(z[i], lo, hi);
                if (i == act) qa = lo;
                if (i < n_out) mx = fmaxf(mx, hi);
            }
            // optim.py:163  targets = r + (1 - d) * gamma * maxQ'
            const float target = meta.x + (1.f - meta.y) * gamma * mx;
            const float d = qa - target;
            if (valid) acc = fmaf(d, d, acc);
            if ((it & (kRegSegLen / 32 - 1)) == kRegSegLen / 32 - 1 || it == n_it - 1) {
                double s = (double)acc;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (__any_sync(0xffffffffu, bad_idx)) s = (double)NAN;
                if (whole) total += s;
                else if (lane == 0) __stcg(part + seg, s);
                acc = 0.f;
                ++seg;
            }
        }
        if (lane == 0) {
            if (whole) {
                fitness.store(cand, (float)(-total / (double)tv.n_trans));
            } else {
                // this unit's segment sums are published; count arrivals with a RELEASE atomic (orders the stores
                // before the increment without the L1 invalidate of a full fence); only the last arriver pays an
                // acquire fence before it reads the other units' sums from L2.
                unsigned prev;
                asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(arrivals + (cand - wu.n_full)) : "memory");
                if (prev == (unsigned)(cur.n_arrive - 1)) {
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    double sum = 0.0;
                    for (int c = 0; c < wu.n_segs; ++c) sum += __ldcg(part + c);   // segment order, as a whole unit would
                    fitness.store(cand, (float)(-sum / (double)tv.n_trans));
                    arrivals[cand - wu.n_full] = 0u;   // self-resetting for the next launch on this stream
                }
            }
        }
        cur = decode(u);
        if (cur.live) build(cur.cand);
    }
}

// Forward only: q[p][b][i]; a thread packs input rows (2j, 2j+1).
template <int N, int LC, int ENT, int NOUT>
__global__ void __launch_bounds__(kRegThreads, 4)
reg_qvalues_kernel(oqrl_spec sp, const float *__restrict__ params, const float *__restrict__ x, int n_rows,
                   float *__restrict__ q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Su2Scalar *coef = reinterpret_cast<Su2Scalar *>(smem_raw);
    const int cand = blockIdx.x;
    const int n_gates = sp.n_layers * N;
    float2 *scr = reinterpret_cast<float2 *>(coef + (sp.n_layers + 1) * N);
    const float *pc = params + (size_t)cand * n_gates * 3;
    build_tables<N, LC, ENT, NOUT>(coef, scr, n_gates, (int)threadIdx.x, (int)blockDim.x,
                                   [&](int j, int) { return angle_of(angle_operands(pc, j), j); }, [] { __syncthreads(); });
    __syncthreads();

    bool emb[N];
    int feat[N];
#pragma unroll
    for (int w = 0; w < N; ++w) {
        feat[w] = embed_feature(w, sp.n_feat, sp.feature_map);
        emb[w] = feat[w] >= 0;
    }
    // sp.reserved bit 0 (set by tests/bench only) disables the light-cone pruning of the last layer
    const unsigned last_mask = (sp.reserved & 1) ? 0xffffffffu : readout_wire_mask(N, sp.n_layers, sp.n_out, ENT);
    const int n_pairs = (n_rows + 1) >> 1;
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < n_pairs; j += gridDim.y * blockDim.x) {
        const int b0 = 2 * j, b1 = min(2 * j + 1, n_rows - 1);
        u64 cw[N], sw[N];
#pragma unroll
        for (int w = 0; w < N; ++w) {
            if (emb[w]) {
                float s0, c0, s1, c1;
                sincosf(0.5f * x[(size_t)b0 * sp.n_feat + feat[w]], &s0, &c0);
                sincosf(0.5f * x[(size_t)b1 * sp.n_feat + feat[w]], &s1, &c1);
                cw[w] = pack2(c0, c1);
                sw[w] = pack2(s0, s1);
            } else {
                cw[w] = pack2(1.f, 1.f);
                sw[w] = 0ull;
            }
        }
        u64 z[N];
        run_circuit<N, LC, ENT, NOUT>(coef, sp.n_layers, sp.reupload, sp.n_out, last_mask, cw, sw, emb, z);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i < sp.n_out) {
                float lo, hi;
                unpack2(z[i], lo, hi);
                q[((size_t)cand * n_rows + b0) * sp.n_out + i] = lo;
                if (b1 != b0) q[((size_t)cand * n_rows + b1) * sp.n_out + i] = hi;
            }
        }
    }
}

// Shared memory of one table owner (a warp of the fitness kernel, the CTA of the forward kernel): coefficient sets of
// every gate + the rotated read-out observables, then the parked half-angle sincos.
inline size_t reg_table_bytes(const oqrl_spec &sp)
{
    return ((size_t)(sp.n_layers + 1) * sp.n_qubits * sizeof(Su2Scalar) + (size_t)3 * sp.n_layers * sp.n_qubits * sizeof(float2) + 15) & ~(size_t)15;
}

// Unit list of one launch for `n_warps` persistent warps.  Whole candidates for as many full rounds as the population
// holds.  A last round that fills at least half of the warps stays whole too (the static deal spreads it evenly over
// the SM sub-partitions, whose throughput does not need all four warps); a smaller one is cut so that every warp gets
// about the same number of segments from at most two kinds of unit (seg_big segments, and what is left of the
// candidate), big ones first.
inline RegUnits plan_units(int n_cand, int n_trans, int n_warps, bool can_split, long long partial_capacity)
{
    RegUnits wu{};
    wu.n_segs = (n_trans + kRegSegLen - 1) / kRegSegLen;
    wu.n_full = n_cand;
    wu.big_per_cand = 1;
    wu.seg_big = wu.n_segs;
    const int rest = n_cand % n_warps;
    if (can_split && wu.n_segs > 1 && rest > 0 && 2 * rest < n_warps && (long long)rest * wu.n_segs <= partial_capacity) {
        const long long share = ((long long)rest * wu.n_segs + n_warps - 1) / n_warps;   // segments per warp in the last round
        if (share < wu.n_segs) {
            wu.n_full = n_cand - rest;
            wu.seg_big = (int)share;
            wu.big_per_cand = wu.n_segs / wu.seg_big;
            wu.seg_rest = wu.n_segs - wu.big_per_cand * wu.seg_big;
            wu.n_big = rest * wu.big_per_cand;
        }
    }
    wu.n_units = wu.n_full + wu.n_big + (wu.seg_rest > 0 ? n_cand - wu.n_full : 0);
    return wu;
}

template <int N, int LC, int ENT, int NOUT, bool STAGED>
int launch_fitness_cfg(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                       const FitnessOut &out, double *d_partial, unsigned *d_arrivals, long long partial_capacity,
                       int sm_count, size_t smem, cudaStream_t st)
{
    constexpr int kWarps = RegCfg<NOUT, STAGED>::kWarps;
    auto kern = reg_fitness_kernel<N, LC, ENT, NOUT, STAGED>;
    // CTAs per SM the register allocation really admits (direct loads; a staged launch is one CTA per SM); the staged
    // kernel's shared-memory opt-in is raised once per instantiation
    static const int resident = [&]() {
        if (STAGED) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRegStageSmemMax) != cudaSuccess) { cudaGetLastError(); return 0; }
            return 1;
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * kWarps, 2048) == cudaSuccess && occ > 0) return occ;
        cudaGetLastError();
        return 3;
    }();
    if (resident == 0) { set_error("reg_fitness: shared-memory opt-in of %d bytes refused", kRegStageSmemMax); return OQRL_ERR_UNSUPPORTED; }
    const int n_warps = resident * sm_count * kWarps;
    const RegUnits wu = plan_units(n_cand, tv.n_trans, n_warps, d_partial && d_arrivals, partial_capacity);
    if (wu.n_units <= 0) return OQRL_OK;
    const int grid = (int)std::min<long long>((long long)resident * sm_count, ((long long)wu.n_units + kWarps - 1) / kWarps);
    kern<<<(unsigned)grid, 32 * kWarps, smem, st>>>(sp, d_params, tv, gamma, wu, out, d_partial, d_arrivals);
    OQRL_LAUNCH_CHECK("reg_fitness_kernel");
    return OQRL_OK;
}

template <int N, int LC, int ENT, int NOUT>
int launch_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                   const FitnessOut &out, double *d_partial, unsigned *d_arrivals, long long partial_capacity,
                   int sm_count, cudaStream_t st)
{
    // the batch is staged in shared memory when it fits beside the tables (sp.reserved bit 2, tests/bench: never)
    const size_t staged = RegCfg<NOUT, true>::kWarps * reg_table_bytes(sp) + (size_t)((tv.n_trans + 31) / 32 + 1) * (1 + N) * 32 * sizeof(float4);
    if (staged <= (size_t)kRegStageSmemMax && !(sp.reserved & 4))
        return launch_fitness_cfg<N, LC, ENT, NOUT, true>(sp, d_params, n_cand, tv, gamma, out, d_partial, d_arrivals,
                                                          partial_capacity, sm_count, staged, st);
    return launch_fitness_cfg<N, LC, ENT, NOUT, false>(sp, d_params, n_cand, tv, gamma, out, d_partial, d_arrivals,
                                                       partial_capacity, sm_count,
                                                       RegCfg<NOUT, false>::kWarps * reg_table_bytes(sp), st);
}

template <int N, int LC, int ENT, int NOUT>
int launch_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q,
                   cudaStream_t st)
{
    const int n_pairs = (n_rows + 1) / 2;
    const int threads = n_pairs >= kRegThreads ? kRegThreads : ((n_pairs + 31) / 32) * 32;
    int gy = (n_pairs + threads - 1) / threads;
    if (gy > 1024) gy = 1024;
    dim3 grid(n_cand, gy);
    reg_qvalues_kernel<N, LC, ENT, NOUT><<<grid, threads, reg_table_bytes(sp), st>>>(sp, d_params, d_x, n_rows, d_q);
    OQRL_LAUNCH_CHECK("reg_qvalues_kernel");
    return OQRL_OK;
}

// Dispatch over (n, unrolled layer count, entangler, compile-time read-out).  The reference circuit
// (4 qubits, 2 layers, CNOT ring, 2 actions, single embedding) gets the fully specialised kernel.
#define OQRL_REG_DISPATCH(FN, ...)                                                                   \
    do {                                                                                             \
        const bool cz = sp.entangler == OQRL_ENT_CZ_RING;                                            \
        switch (sp.n_qubits) {                                                                       \
        case 1: return FN<1, 0, OQRL_ENT_SEL_CNOT, 0>(__VA_ARGS__);                                  \
        case 2: return cz ? FN<2, 0, OQRL_ENT_CZ_RING, 0>(__VA_ARGS__) : FN<2, 0, OQRL_ENT_SEL_CNOT, 0>(__VA_ARGS__); \
        case 3: return cz ? FN<3, 0, OQRL_ENT_CZ_RING, 0>(__VA_ARGS__) : FN<3, 0, OQRL_ENT_SEL_CNOT, 0>(__VA_ARGS__); \
        case 4:                                                                                      \
            if (cz && !sp.reupload && sp.n_out == 2 && sp.n_feat == 4) {                             \
                const bool np_ = (sp.reserved & 1) != 0;                                             \
                switch (sp.n_layers) {                                                               \
                case 1: return np_ ? FN<4, 1, OQRL_ENT_CZ_RING, -2>(__VA_ARGS__) : FN<4, 1, OQRL_ENT_CZ_RING, 2>(__VA_ARGS__); \
                case 2: return np_ ? FN<4, 2, OQRL_ENT_CZ_RING, -2>(__VA_ARGS__) : FN<4, 2, OQRL_ENT_CZ_RING, 2>(__VA_ARGS__); \
                case 3: return np_ ? FN<4, 3, OQRL_ENT_CZ_RING, -2>(__VA_ARGS__) : FN<4, 3, OQRL_ENT_CZ_RING, 2>(__VA_ARGS__); \
                case 4: return np_ ? FN<4, 4, OQRL_ENT_CZ_RING, -2>(__VA_ARGS__) : FN<4, 4, OQRL_ENT_CZ_RING, 2>(__VA_ARGS__); \
                default: break;                                                                      \
                }                                                                                    \
            }                                                                                        \
            if (cz) return FN<4, 0, OQRL_ENT_CZ_RING, 0>(__VA_ARGS__);                               \
            if (!sp.reupload && sp.n_out == 2 && sp.n_feat == 4) {                                   \
                /* the reference agent (two actions, four features), layer counts unrolled */        \
                const bool np_ = (sp.reserved & 1) != 0;                                             \
                switch (sp.n_layers) {                                                               \
                case 1: return np_ ? FN<4, 1, OQRL_ENT_SEL_CNOT, -2>(__VA_ARGS__) : FN<4, 1, OQRL_ENT_SEL_CNOT, 2>(__VA_ARGS__); \
                case 2: return np_ ? FN<4, 2, OQRL_ENT_SEL_CNOT, -2>(__VA_ARGS__) : FN<4, 2, OQRL_ENT_SEL_CNOT, 2>(__VA_ARGS__); \
                case 3: return np_ ? FN<4, 3, OQRL_ENT_SEL_CNOT, -2>(__VA_ARGS__) : FN<4, 3, OQRL_ENT_SEL_CNOT, 2>(__VA_ARGS__); \
                case 4: return np_ ? FN<4, 4, OQRL_ENT_SEL_CNOT, -2>(__VA_ARGS__) : FN<4, 4, OQRL_ENT_SEL_CNOT, 2>(__VA_ARGS__); \
                default: break;                                                                      \
                }                                                                                    \
            }                                                                                        \
            return FN<4, 0, OQRL_ENT_SEL_CNOT, 0>(__VA_ARGS__);                                      \
        default: break;                                                                              \
        }                                                                                            \
    } while (0)

}  // namespace

int reg_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                const FitnessOut &out, double *d_partial, unsigned *d_arrivals, long long partial_capacity,
                int sm_count, cudaStream_t st)
{
    if (sp.n_layers * sp.n_qubits > kRegMaxGates) {
        set_error("register-resident family supports n_layers*n_qubits <= %d", kRegMaxGates);
        return OQRL_ERR_UNSUPPORTED;
    }
    OQRL_REG_DISPATCH(launch_fitness, sp, d_params, n_cand, tv, gamma, out, d_partial, d_arrivals, partial_capacity, sm_count, st);
    set_error("reg_fitness: n_qubits=%d outside 1..4", sp.n_qubits);
    return OQRL_ERR_INVALID;
}

int reg_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q,
                cudaStream_t st)
{
    if (sp.n_layers * sp.n_qubits > kRegMaxGates) {
        set_error("register-resident family supports n_layers*n_qubits <= %d", kRegMaxGates);
        return OQRL_ERR_UNSUPPORTED;
    }
    OQRL_REG_DISPATCH(launch_qvalues, sp, d_params, n_cand, d_x, n_rows, d_q, st);
    set_error("reg_qvalues: n_qubits=%d outside 1..4", sp.n_qubits);
    return OQRL_ERR_INVALID;
}

// Test hook (exported by the probe library as oqrl_reg_plan): the unit list plan_units deals to `n_warps` persistent
// warps, as seven int32 {n_full, n_big, big_per_cand, seg_big, seg_rest, n_units, n_segs}.
void reg_plan_dump(int n_cand, int n_trans, int n_warps, int can_split, long long partial_capacity, int32_t *out)
{
    const RegUnits wu = plan_units(n_cand, n_trans, n_warps, can_split != 0, partial_capacity);
    out[0] = wu.n_full; out[1] = wu.n_big; out[2] = wu.big_per_cand; out[3] = wu.seg_big; out[4] = wu.seg_rest;
    out[5] = wu.n_units; out[6] = wu.n_segs;
}

// Executed fp32 flop per circuit evaluation (each half of a packed op counts separately).
double reg_flops_per_eval(const oqrl_spec &sp)
{
    const int n = sp.n_qubits, L = sp.n_layers;
    const double dim = (double)(1 << n);
    int n_emb = 0;
    for (int w = 0; w < n; ++w) n_emb += embed_feature(w, sp.n_feat, sp.feature_map) >= 0;
    double f;
    // the fully specialised reference circuit (OQRL_REG_DISPATCH) builds its product state as a balanced tree
    const bool specialised = n == 4 && L >= 1 && L <= 4 && !sp.reupload && sp.n_out == 2 && sp.n_feat == 4;
    if (L > 0) f = 12.0 * n + 6.0 * (specialised ? 24.0 : 2.0 * dim - 4.0);   // per-wire Rot on a 2-vector + complex tensor product
    else f = 2.0 * dim - 4.0;                            // real tensor product
    // ... and folds its last layer into rotated observables
    const bool rotated = specialised && !(sp.reserved & 1) && readout_is_single_wire(n, L, sp.n_out, sp.entangler);
    if (rotated) {
        f += 14.0 * dim * n * (L - 2);                   // layers 1 .. L-2 in full; the last one lives in the observables
        // |psi|^2 sector sums shared by the outputs (2 dim FMAs, one leading multiply per sector); per output the C_r chain
        // (dim FMAs, first op a multiply), the two halves of C_i (dim/2 each) and the (sectors + 3)-term combine
        const double sectors = (double)(1 << sp.n_out);
        f += 4.0 * dim - sectors;
        f += (double)sp.n_out * ((2.0 * dim - 1.0) + 2.0 * (dim - 1.0) + 2.0 * (sectors + 3.0) - 1.0);
        return f;
    }
    if (L > 1) {
        const unsigned lm = (sp.reserved & 1) ? 0xffffffffu : readout_wire_mask(n, L, sp.n_out, sp.entangler);
        int last_gates = 0, last_emb = 0;
        for (int w = 0; w < n; ++w)
            if ((lm >> w) & 1u) { ++last_gates; last_emb += embed_feature(w, sp.n_feat, sp.feature_map) >= 0; }
        f += 14.0 * dim * (n * (L - 2) + last_gates);    // SU(2) gate: 28 flop per amplitude pair; last layer pruned
        if (sp.reupload) f += 6.0 * dim * (n_emb * (L - 2) + last_emb);   // RY: 12 flop per pair
    }
    f += 3.0 * dim + (double)sp.n_out * dim;             // |amp|^2 (3 flop) + n_out signed sums
    return f;
}

}  // namespace oqrl
