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
// Z-string is one wire w, and  <Z> after its gate U = [[a, b], [-conj b, conj a]]  is
//     n_z (S0 - S1) + G_r C_r + G_i C_i,   S0/S1 = sum |psi|^2 over the wire's 0/1 half,  C = sum conj(psi_0) psi_1,
//     n_z = |a|^2 - |b|^2,  G_r = 4 Re(conj(a) b),  G_i = -4 Im(conj(a) b)            (obs[i] = {n_z, G_r, G_i, -}).
// Five accumulate-in-place FMA chains per output (64 packed FMAs) replace the gate (128) and the |amp|^2 read-out.
template <int N, int LC, int ENT, int NOUT>
__device__ __forceinline__ void readout_rotated(const RegState<N> &st, const Su2Scalar *__restrict__ obs, u64 (&z)[N])
{
#pragma unroll
    for (int i = 0; i < N; ++i) {
        z[i] = 0ull;
        if (i >= NOUT) continue;
        constexpr unsigned kAll = (1u << N) - 1u;
        const unsigned m = readout_index_mask(N, LC, i, ENT) & kAll;
        int pos = 0;
#pragma unroll
        for (int b = 0; b < N; ++b)
            if ((m >> b) & 1u) pos = b;
        u64 s0 = 0ull, s1 = 0ull, cr = 0ull, ca = 0ull, cb = 0ull;
        bool first = true;
#pragma unroll
        for (int k = 0; k < (1 << N); ++k) {
            if ((k >> pos) & 1) continue;
            const int k1 = k | (1 << pos);
            if (first) {
                s0 = fmul2(st.re[k], st.re[k]);
                s1 = fmul2(st.re[k1], st.re[k1]);
                cr = fmul2(st.re[k], st.re[k1]);
                ca = fmul2(st.re[k], st.im[k1]);
                cb = fmul2(st.im[k], st.re[k1]);
                first = false;
            } else {
                s0 = ffma2(st.re[k], st.re[k], s0);
                s1 = ffma2(st.re[k1], st.re[k1], s1);
                cr = ffma2(st.re[k], st.re[k1], cr);
                ca = ffma2(st.re[k], st.im[k1], ca);
                cb = ffma2(st.im[k], st.re[k1], cb);
            }
            s0 = ffma2(st.im[k], st.im[k], s0);
            s1 = ffma2(st.im[k1], st.im[k1], s1);
            cr = ffma2(st.im[k], st.im[k1], cr);
        }
        const Su2Scalar o = obs[i];
        u64 q = fmul2(bcast2(o.ar), s0);
        q = ffma2(bcast2(-o.ar), s1, q);
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

// Prologue: candidate's Rot gates -> shared memory, pre-duplicated for fp32x2.
__device__ __forceinline__ void build_gate_table(Su2Scalar *coef, const float *__restrict__ params, int n_gates)
{
    for (int g = threadIdx.x; g < n_gates; g += blockDim.x) {
        Su2Scalar c;
        rot_su2(params[3 * g + 0], params[3 * g + 1], params[3 * g + 2], c.ar, c.ai, c.br, c.bi);
        coef[g] = c;
    }
}

// Prologue of the specialised kernels: gate table plus, appended to it, the rotated read-out observables
// (readout_rotated) -- one pass, so the observables cost no second round of sincos.
template <int N, int LC, int ENT, int NOUT>
__device__ __forceinline__ void build_tables(Su2Scalar *coef, const float *__restrict__ params, int n_gates)
{
    constexpr bool kRotated = NOUT > 0 && LC >= 2 && readout_is_single_wire(N, LC, NOUT > 0 ? NOUT : 1, ENT);
    if constexpr (!kRotated) {
        build_gate_table(coef, params, n_gates);
    } else {
        for (int e = threadIdx.x; e < LC * N + NOUT; e += blockDim.x) {
            int g = e;
            if (e >= LC * N) {
                g = 0;
#pragma unroll
                for (int i = 0; i < NOUT; ++i) {
                    const unsigned m = readout_index_mask(N, LC, i, ENT);
                    int pos = 0;
#pragma unroll
                    for (int b = 0; b < N; ++b)
                        if ((m >> b) & 1u) pos = b;
                    if (e - LC * N == i) g = (LC - 1) * N + (N - 1 - pos);
                }
            }
            float ar, ai, br, bi;
            rot_su2(params[3 * g + 0], params[3 * g + 1], params[3 * g + 2], ar, ai, br, bi);
            Su2Scalar c;
            if (e >= LC * N) {
                c.ar = ar * ar + ai * ai - br * br - bi * bi;
                c.ai = 4.f * (ar * br + ai * bi);
                c.br = -4.f * (ar * bi - ai * br);
                c.bi = 0.f;
            } else {
                c.ar = ar; c.ai = ai; c.br = br; c.bi = bi;
            }
            coef[e] = c;
        }
    }
}

// One WARP (= one 32-thread CTA) per work item (candidate, chunk of transitions): no CTA barrier anywhere,
// so warps never wait for one another (the 128-thread/one-barrier version lost 11% of its issue slots to
// barrier stalls -- profiles/r1_reg_fitness_v1.md).  With several chunks per candidate the per-chunk sums
// meet in `partial`; the last chunk to arrive adds them in chunk order (deterministic) and writes the fitness.
// 12 -> ptxas takes 138 registers (14 resident warps per SM): measured 1.6 % faster than the 128-register / 16-warp build
#ifndef OQRL_REG_MIN_WARPS
#define OQRL_REG_MIN_WARPS 12
#endif
template <int N, int LC, int ENT, int NOUT>
__global__ void __launch_bounds__(32, OQRL_REG_MIN_WARPS)
reg_fitness_kernel(oqrl_spec sp, const float *__restrict__ params, TransitionView tv, float gamma,
                   int chunk_len, int n_chunks_split, int n_full, FitnessOut fitness, double *__restrict__ partial,
                   unsigned *__restrict__ arrivals)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Su2Scalar *coef = reinterpret_cast<Su2Scalar *>(smem_raw);

    // The first n_full work items are whole candidates (all transitions, no chunk protocol, one table build per 32+
    // iterations); the rest of the population is split into n_chunks_split chunks per candidate.  Long items are
    // launched first, so the hardware CTA scheduler fills the tail of the grid with the short ones.
    const bool whole = (int)blockIdx.x < n_full;
    const int n_chunks = whole ? 1 : n_chunks_split;
    const int rest = whole ? 0 : (int)blockIdx.x - n_full;
    const int cand = whole ? (int)blockIdx.x : n_full + rest / n_chunks_split;
    const int chunk = whole ? 0 : rest - (cand - n_full) * n_chunks_split;
    const int n_gates = sp.n_layers * N;
    build_tables<N, LC, ENT, NOUT>(coef, params + (size_t)cand * n_gates * 3, n_gates);
    __syncwarp();

    bool emb[N];
    int feat[N];
#pragma unroll
    for (int w = 0; w < N; ++w) {
        feat[w] = embed_feature(w, sp.n_feat, sp.feature_map);
        emb[w] = feat[w] >= 0;
    }
    // sp.reserved bit 0 (set by tests/bench only) disables the light-cone pruning of the last layer
    const unsigned last_mask = (sp.reserved & 1) ? 0xffffffffu : readout_wire_mask(N, sp.n_layers, sp.n_out, ENT);

    const int t0 = chunk * chunk_len;
    const int t1 = whole ? tv.n_trans : min(tv.n_trans, t0 + chunk_len);
    float acc = 0.f;
    // Software pipeline: the half-angle record of iteration i+1 is loaded (into the same registers) right after
    // the product state of iteration i has consumed the current one, the row index of iteration i+2 likewise,
    // so the dependent idx -> row -> record chain never stalls the warp; the TD metadata is only needed at the
    // end of an iteration and is loaded at its top.
    int t = t0 + (int)threadIdx.x;
    // (a loaded index is not touched until the next iteration: no select on it, the bound check is on `tt`)
    auto load_row = [&](int tt) { const int tc = min(tt, t1 - 1); return tv.idx ? __ldg(tv.idx + tc) : tc; };
    float4 trig_reg[N];
    // NOUT != 0 marks the fully specialised reference circuit (|NOUT| outputs; NOUT < 0 = without pruning): every wire embedded, feature index = wire
    // (the dispatcher checks n_feat == N), so the record is N consecutive float4 and nothing is predicated.
    constexpr bool kAllEmb = NOUT != 0;
    auto load_trig = [&](int row) {
        const float4 *trig = tv.trig + (size_t)row * (kAllEmb ? N : sp.n_feat);
#pragma unroll
        for (int w = 0; w < N; ++w)
            if (kAllEmb || emb[w]) trig_reg[w] = __ldg(trig + (kAllEmb ? w : feat[w]));
    };
    bool bad_idx = false;                // an index outside the dataset: NaN fitness (checked where the index is consumed)
    int row_cur = checked_row(load_row(t), tv.n_rows, bad_idx);   // clamped: lanes past the end read a valid row and discard the result
    load_trig(row_cur);
    int row_next = load_row(t + 32);
    // The transitions of a generation are cut into fixed SEGMENTS of kRegSegLen.  Every segment is reduced the same
    // way (fp32 per lane, fp64 shuffle tree) and a candidate's fitness is the fp64 sum of its segment sums in segment
    // order -- whether one work item ran all of them or several items ran a few each.  So the bits of a fitness value
    // do not depend on the population size, on the candidate's position in it, or on how it is sharded over GPUs.
    // (A split work item is exactly one segment: chunk_len == kRegSegLen whenever n_chunks_split > 1.)
    double total = 0.0;
    int seg_end = t0 + chunk_len;
    for (int tb = t0; tb < t1; tb += 32, t += 32) {      // warp-uniform trip count
        const bool valid = t < t1;
        const float4 meta = __ldg(tv.meta + row_cur);
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
            row_next = checked_row(row_next, tv.n_rows, bad_idx);   // loaded one iteration ago: no stall
            if (t + 32 < t1) load_trig(row_next);
            row_cur = row_next;
            row_next = load_row(t + 64);
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
        if (tb + 32 >= seg_end || tb + 32 >= t1) {
            double s = (double)acc;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            total += s;
            acc = 0.f;
            seg_end += chunk_len;
        }
    }
    const double tot = __any_sync(0xffffffffu, bad_idx) ? (double)NAN : total;
    if (threadIdx.x == 0) {
        if (n_chunks == 1) {
            fitness.store(cand, (float)(-tot / (double)tv.n_trans));
        } else {
            // publish this segment's sum, then count arrivals with a RELEASE atomic (orders the store before
            // the increment without the L1 invalidate of a full fence); only the last arriver pays an
            // acquire fence before it reads the other segments' sums from L2.
            double *part = partial + (size_t)(cand - n_full) * n_chunks;
            __stcg(part + chunk, tot);
            unsigned prev;
            asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(arrivals + (cand - n_full)) : "memory");
            if (prev == (unsigned)(n_chunks - 1)) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                double sum = 0.0;
                for (int c = 0; c < n_chunks; ++c) sum += __ldcg(part + c);   // segment order, as a whole item would
                fitness.store(cand, (float)(-sum / (double)tv.n_trans));
                arrivals[cand - n_full] = 0u;   // self-resetting for the next launch on this stream
            }
        }
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
    build_tables<N, LC, ENT, NOUT>(coef, params + (size_t)cand * n_gates * 3, n_gates);
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

// Work items = candidates x chunks.  Chunks are cut so that (a) the grid is several waves of the resident
// warps per SM, and (b) every lane still runs a few transitions per gate-table build.
inline int pick_chunks(int n_cand, int n_trans, int sm_count)
{
    const int iters = (n_trans + 31) / 32;                 // lane-iterations per candidate
    const long long want_items = 6LL * 16 * sm_count;
    long long chunks = (want_items + n_cand - 1) / n_cand;
    const int max_chunks = iters >= 8 ? iters / 8 : 1;     // keep >= 8 iterations per chunk
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    // prefer an even split of the iterations (no short tail chunk) if one exists within 2x
    for (long long c = chunks; c <= 2 * chunks && c <= max_chunks; ++c)
        if (iters % c == 0) { chunks = c; break; }
    return (int)chunks;
}

template <int N, int LC, int ENT, int NOUT>
int launch_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                   const FitnessOut &out, double *d_partial, unsigned *d_arrivals, long long partial_capacity, int sm_count,
                   cudaStream_t st)
{
    const int T = tv.n_trans;
    // whole candidates for as many full waves of resident warps as the population holds ...
    // warps (= CTAs) per SM the register allocation really admits; queried once per kernel instantiation
    static const int resident = []() {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reg_fitness_kernel<N, LC, ENT, NOUT>, 32, 1024) == cudaSuccess && occ > 0) return occ;
        cudaGetLastError();
        return (int)OQRL_REG_MIN_WARPS;
    }();
    const int slots = resident * sm_count;
    int n_full = (d_partial && d_arrivals && T >= 256) ? (n_cand / slots) * slots : 0;
    // a last wave that is (nearly) full needs no splitting at all
    if (n_full > 0 && (n_cand - n_full) * 8 >= slots * 7) n_full = n_cand;
    const int n_split = n_cand - n_full;
    // ... and the remainder cut into chunks that fill the last partial wave
    // ... and the remainder split into work items of exactly one segment each (the scratch holds one fp64 sum per
    // segment and split candidate).  Whole items reduce the same segments in the same order, so the two forms agree
    // bit for bit and a fitness value never depends on the launch shape.
    const long long n_segs = (T + kRegSegLen - 1) / kRegSegLen;
    int chunks = (n_split > 0 && d_partial && d_arrivals && pick_chunks(n_split + (n_full > 0 ? slots : 0), T, sm_count) > 1) ? (int)n_segs : 1;
    if (chunks > 1 && (long long)n_split * n_segs > partial_capacity) chunks = 1;
    const int chunk_len = kRegSegLen;     // segment length of whole items == length of a split item
    if (chunks == 1) n_full = n_cand;     // nothing to split: every candidate is a whole work item
    const long long items = (long long)n_full + (long long)(n_cand - n_full) * chunks;
    if (items > 0x7fffffffLL) { set_error("grid too large"); return OQRL_ERR_INVALID; }
    const size_t smem = (size_t)(sp.n_layers + 1) * N * sizeof(Su2Scalar);   // gates + rotated read-out observables
    reg_fitness_kernel<N, LC, ENT, NOUT><<<(unsigned)items, 32, smem, st>>>(sp, d_params, tv, gamma, chunk_len, chunks, n_full,
                                                                           out, d_partial, d_arrivals);
    OQRL_LAUNCH_CHECK("reg_fitness_kernel");
    return OQRL_OK;
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
    reg_qvalues_kernel<N, LC, ENT, NOUT><<<grid, threads, (size_t)(sp.n_layers + 1) * N * sizeof(Su2Scalar), st>>>(sp, d_params, d_x, n_rows, d_q);
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
                const FitnessOut &out, double *d_partial, unsigned *d_arrivals, long long partial_capacity, int sm_count, cudaStream_t st)
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
        // per output: S0, S1, C_r chains (dim FMAs each, first op a multiply), the two halves of C_i (dim/2 each), 5-term combine
        f += (double)sp.n_out * (3.0 * (2.0 * dim - 1.0) + 2.0 * (dim - 1.0) + 9.0);
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
