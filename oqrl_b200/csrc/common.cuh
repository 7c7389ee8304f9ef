// Shared declarations for the oqrl_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/oqrl_b200.h"

namespace oqrl {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define OQRL_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::oqrl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                              __FILE__, __LINE__);                                         \
            return OQRL_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define OQRL_LAUNCH_CHECK(name)                                                            \
    do {                                                                                   \
        ::oqrl::count_launch();                                                            \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            ::oqrl::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
            return OQRL_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

// ---- circuit structure helpers (host + device) --------------------------------
// StronglyEntanglingLayers default range r_l = (l mod (n-1)) + 1.
__host__ __device__ constexpr int sel_range(int n, int layer) { return n > 1 ? (layer % (n - 1)) + 1 : 0; }

// Feature index rotated onto `wire` by the embedding, or -1 if the wire is not embedded.
__host__ __device__ inline int embed_feature(int wire, int n_feat, int feature_map)
{
    if (n_feat <= 0) return -1;
    if (feature_map == OQRL_FEAT_TILED) return wire % n_feat;
    return wire < n_feat ? wire : -1;
}

// Basis-label image of the whole CNOT ring of one layer: |k> -> |sel_perm(k)>,
// wire 0 = most significant bit of k.
__host__ __device__ constexpr unsigned sel_perm(unsigned k, int n, int r)
{
    for (int i = 0; i < n; ++i) {
        int t = (i + r) % n;
        if ((k >> (n - 1 - i)) & 1u) k ^= 1u << (n - 1 - t);
    }
    return k;
}

// Light cone of the read-out: wires whose LAST-layer gate can influence <Z_i>, i < n_out.  The final
// entangler only relabels basis states (CNOT ring) or flips signs (CZ ring), so <Z_i> after it is the
// parity of a fixed set of pre-entangler wires; a single-qubit gate on any other wire commutes with the
// measured observable and is skipped (exact, not an approximation).  Bit w of the result = wire w needed.
__host__ __device__ constexpr unsigned readout_wire_mask(int n, int n_layers, int n_out, int entangler)
{
    unsigned mask = 0;
    if (n_layers <= 0) return 0;
    if (n < 2 || entangler != OQRL_ENT_SEL_CNOT) {
        for (int i = 0; i < n_out; ++i) mask |= 1u << i;
        return mask;
    }
    const int r = sel_range(n, n_layers - 1);
    for (int w = 0; w < n; ++w) {
        const unsigned img = sel_perm(1u << (n - 1 - w), n, r);   // which output bits flip when wire w flips
        for (int i = 0; i < n_out; ++i)
            if ((img >> (n - 1 - i)) & 1u) mask |= 1u << w;
    }
    return mask;
}

// <Z_i> measured after the last layer's entangler equals a Z-string on the state BEFORE it: the set of index bits
// whose parity gives the sign (CNOT ring: row of the relabelling; CZ ring / no entangler: the wire's own bit).
__host__ __device__ constexpr unsigned readout_index_mask(int n, int n_layers, int i, int entangler)
{
    if (n_layers <= 0 || n < 2 || entangler != OQRL_ENT_SEL_CNOT) return 1u << (n - 1 - i);
    const int r = sel_range(n, n_layers - 1);
    unsigned m = 0;
    for (int b = 0; b < n; ++b)
        if ((sel_perm(1u << b, n, r) >> (n - 1 - i)) & 1u) m |= 1u << b;
    return m;
}

// True when every read-out's pre-entangler Z-string is a single wire.  Then the last layer needs no gate at all:
// <Z> after the gate U on that wire is the expectation of the rotated observable U^+ Z U on the state before it,
// i.e. n_z <Z_w> + 4 Re(conj(a) b <sigma^-_w>) -- three sums over amplitude pairs (reg kernel, reference circuit).
__host__ __device__ constexpr bool readout_is_single_wire(int n, int n_layers, int n_out, int entangler)
{
    if (n_layers < 2) return false;
    for (int i = 0; i < n_out; ++i) {
        const unsigned m = readout_index_mask(n, n_layers, i, entangler);
        if (m == 0 || (m & (m - 1)) != 0) return false;
    }
    return true;
}

// Sign of basis state k under the CZ ring CZ(i,(i+1) mod n), i = 0..n-1.
__host__ __device__ constexpr bool cz_ring_negative(unsigned k, int n)
{
    int par = 0;
    if (n < 2) return false;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        par ^= (int)((k >> (n - 1 - i)) & (k >> (n - 1 - j)) & 1u);
    }
    return par != 0;
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2) ----------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fsub2(u64 a, u64 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#ifdef OQRL_FNEG2_XOR
__device__ __forceinline__ u64 fneg2(u64 a) { return a ^ 0x8000000080000000ull; }
#else
// written as two float negations: ptxas folds them into the `-R.F32x2` operand modifier of the FFMA2 / FMUL2 that
// consumes the value (no instruction at all); the XOR form costs two LOP3 per negation
__device__ __forceinline__ u64 fneg2(u64 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(-lo), "f"(-hi));
    return r;
}
#endif

// 16-byte shared-memory record <-> two packed values (shared-window address)
__device__ __forceinline__ void st128(unsigned addr, u64 a, u64 b)
{
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld128(unsigned addr, u64 &a, u64 &b)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}

// SU(2) gate [[a, b], [-conj(b), conj(a)]].  Every single-qubit gate on this path (RY, Rot,
// and their products) has this form, so four reals describe it.
//   Su2Coef  : per-circuit coefficients (the two packed halves differ -- re-uploaded gates);
//   Su2Scalar: one coefficient set for both halves (candidate-uniform Rot gates).  Kept as
//              32-bit scalars: FFMA2 broadcasts a scalar operand to both halves (SASS `Rn.F32`),
//              which halves the register footprint and the register-file traffic of the gate.
struct Su2Coef {
    u64 ar, ai, nai, br, nbr, bi, nbi, pad;
};
struct __align__(16) Su2Scalar {
    float ar, ai, br, bi;
};
__device__ __forceinline__ u64 bcast2(float v) { return pack2(v, v); }

// Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi)  (SURVEY.md App. A.3a):
//   a = e^{-i(phi+omega)/2} cos(theta/2),  b = -e^{+i(phi-omega)/2} sin(theta/2)
__device__ __forceinline__ void rot_su2(float phi, float theta, float omega,
                                        float &ar, float &ai, float &br, float &bi)
{
    float c, s, cp, sp, cq, sq;
    sincosf(0.5f * theta, &s, &c);
    sincosf(0.5f * (phi + omega), &sp, &cp);
    sincosf(0.5f * (phi - omega), &sq, &cq);
    ar = cp * c;
    ai = -sp * c;
    br = -cq * s;
    bi = -sq * s;
}

// One amplitude pair (x0 = target bit 0, x1 = target bit 1) through an SU(2) gate:
// 4 FMUL2 + 12 FFMA2 for two circuits at once.
__device__ __forceinline__ void su2_pair(const Su2Coef &g, u64 &x0r, u64 &x0i, u64 &x1r, u64 &x1i)
{
    u64 n0r = fmul2(g.ar, x0r);
    u64 n0i = fmul2(g.ar, x0i);
    u64 n1r = fmul2(g.nbr, x0r);
    u64 n1i = fmul2(g.nbr, x0i);
    n0r = ffma2(g.nai, x0i, n0r);
    n0i = ffma2(g.ai, x0r, n0i);
    n1r = ffma2(g.nbi, x0i, n1r);
    n1i = ffma2(g.bi, x0r, n1i);
    n0r = ffma2(g.br, x1r, n0r);
    n0i = ffma2(g.br, x1i, n0i);
    n1r = ffma2(g.ar, x1r, n1r);
    n1i = ffma2(g.ar, x1i, n1i);
    n0r = ffma2(g.nbi, x1i, n0r);
    n0i = ffma2(g.bi, x1r, n0i);
    n1r = ffma2(g.ai, x1i, n1r);
    n1i = ffma2(g.nai, x1r, n1i);
    x0r = n0r; x0i = n0i; x1r = n1r; x1i = n1i;
}

// Same update with a scalar (both-halves) coefficient set.
__device__ __forceinline__ void su2_pair(const Su2Scalar &g, u64 &x0r, u64 &x0i, u64 &x1r, u64 &x1i)
{
    const u64 ar = bcast2(g.ar), ai = bcast2(g.ai), nai = bcast2(-g.ai);
    const u64 br = bcast2(g.br), nbr = bcast2(-g.br), bi = bcast2(g.bi), nbi = bcast2(-g.bi);
    u64 n0r = fmul2(ar, x0r);
    u64 n0i = fmul2(ar, x0i);
    u64 n1r = fmul2(nbr, x0r);
    u64 n1i = fmul2(nbr, x0i);
    n0r = ffma2(nai, x0i, n0r);
    n0i = ffma2(ai, x0r, n0i);
    n1r = ffma2(nbi, x0i, n1r);
    n1i = ffma2(bi, x0r, n1i);
    n0r = ffma2(br, x1r, n0r);
    n0i = ffma2(br, x1i, n0i);
    n1r = ffma2(ar, x1r, n1r);
    n1i = ffma2(ar, x1i, n1i);
    n0r = ffma2(nbi, x1i, n0r);
    n0i = ffma2(bi, x1r, n0i);
    n1r = ffma2(ai, x1i, n1r);
    n1i = ffma2(nai, x1r, n1i);
    x0r = n0r; x0i = n0i; x1r = n1r; x1i = n1i;
}

// Real rotation RY: a = (c, 0), b = (-s, 0): 4 FMUL2 + 4 FFMA2 per pair.
__device__ __forceinline__ void ry_pair(u64 c, u64 s, u64 ns, u64 &x0r, u64 &x0i, u64 &x1r, u64 &x1i)
{
    u64 n0r = fmul2(c, x0r);
    u64 n0i = fmul2(c, x0i);
    u64 n1r = fmul2(s, x0r);
    u64 n1i = fmul2(s, x0i);
    n0r = ffma2(ns, x1r, n0r);
    n0i = ffma2(ns, x1i, n0i);
    n1r = ffma2(c, x1r, n1r);
    n1i = ffma2(c, x1i, n1i);
    x0r = n0r; x0i = n0i; x1r = n1r; x1i = n1i;
}

// ---- complex arithmetic on packed (re, im) -----------------------------------------------------------------------
__device__ __forceinline__ u64 cswap(u64 v)          // (re, im) -> (im, re): folded into FFMA2's .LO_HI operand form
{
    float lo, hi;
    unpack2(v, lo, hi);
    return pack2(hi, lo);
}
__device__ __forceinline__ u64 cmulp(u64 a, u64 b)   // complex product: 1 FMUL2 + 1 FFMA2
{
    float ar, ai;
    unpack2(a, ar, ai);
    return ffma2(pack2(-ai, ai), cswap(b), fmul2(pack2(ar, ar), b));
}
// SU(2) gate [[a, b], [-conj b, conj a]] on the amplitude pair (x0, x1), amplitudes packed (re, im):
//   n0 = a x0 + b x1,  n1 = -conj(b) x0 + conj(a) x1     -- 2 FMUL2 + 6 FFMA2
// Real parts multiply as broadcast scalars (SASS `R.F32`); the imaginary parts multiply the half-swapped amplitude
// with one half negated, which FFMA2 takes as operand modifiers of a register PAIR (ai, ai).  The coefficient table
// therefore stores those two pairs duplicated and they are LOADED as 64-bit values: a pair ptxas can see through
// (mov.b64 {v, v}) is rebuilt with two moves in front of every use -- 212 moves per 256 FFMA2 in the first build.
struct GateOps {
    float ar, br;
    u64 pai, pbi;    // (ai, ai), (bi, bi)
};
__device__ __forceinline__ u64 neg_lo(u64 v) { float lo, hi; unpack2(v, lo, hi); return pack2(-lo, hi); }
__device__ __forceinline__ u64 neg_hi(u64 v) { float lo, hi; unpack2(v, lo, hi); return pack2(lo, -hi); }
__device__ __forceinline__ GateOps load_gate_ops(const float4 *rec)
{
    GateOps g;
    const float2 re = *reinterpret_cast<const float2 *>(rec);
    const ulonglong2 im = *reinterpret_cast<const ulonglong2 *>(rec + 1);
    g.ar = re.x; g.br = re.y;
    g.pai = im.x; g.pbi = im.y;
    return g;
}
__device__ __forceinline__ void su2_cpair(const GateOps &g, u64 &x0, u64 &x1)
{
    const u64 a1 = pack2(g.ar, g.ar), b1 = pack2(g.br, g.br), b1n = pack2(-g.br, -g.br);
    const u64 s0 = cswap(x0), s1 = cswap(x1);
    u64 n0 = fmul2(a1, x0);
    u64 n1 = fmul2(b1n, x0);
    n0 = ffma2(neg_lo(g.pai), s0, n0);
    n1 = ffma2(neg_lo(g.pbi), s0, n1);
    n0 = ffma2(b1, x1, n0);
    n1 = ffma2(a1, x1, n1);
    n0 = ffma2(neg_lo(g.pbi), s1, n0);
    n1 = ffma2(neg_hi(g.pai), s1, n1);
    x0 = n0; x1 = n1;
}

// ---- kernel-family entry points (one .cu each) ----------------------------------
struct TransitionView {
    // Half-angle table: for transition row j and feature f,
    //   trig[(j*n_feat + f)*4 + {0,1,2,3}] = cos(s/2), cos(s'/2), sin(s/2), sin(s'/2)
    const float4 *trig;
    // meta[j] = {reward, terminal, action (int bits), 0}
    const float4 *meta;
    const int32_t *idx;  // row of transition t, or NULL for identity
    int32_t n_trans;
    int32_t n_rows;      // rows behind trig/meta: an idx entry outside [0, n_rows) poisons every fitness value (NaN)
};

// Range check of a gathered row index, uniform across kernel families (the reference's fancy indexing raises
// IndexError, dataset.py:60-66): the row is replaced by 0 so nothing is read out of bounds and `bad` is raised;
// a launch that saw a bad index writes NaN fitness for the candidates it covered.
__device__ __forceinline__ int checked_row(int row, int n_rows, bool &bad)
{
    if ((unsigned)row >= (unsigned)n_rows) { bad = true; row = 0; }
    return row;
}

// Where a fitness value goes: the local vector, or (fused all-gather) the full-population vector of every peer.
constexpr int kMaxPeers = 16;
struct FitnessOut {
    float *ptr[kMaxPeers];   // ptr[0] only when n == 1
    int n;                   // number of destination buffers
    int offset;              // index of local candidate 0 inside the destination buffers
    __device__ __forceinline__ void store(int cand, float v) const
    {
#pragma unroll 1
        for (int r = 0; r < n; ++r) ptr[r][offset + cand] = v;
    }
};

// register-resident family (n <= 4): pqc_reg.cu
int reg_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv,
                float gamma, const FitnessOut &out, double *d_partial, unsigned *d_arrivals, long long partial_capacity,
                int sm_count, cudaStream_t st);
int reg_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows,
                float *d_q, cudaStream_t st);
double reg_flops_per_eval(const oqrl_spec &sp);
void reg_plan_dump(int n_cand, int n_trans, int n_warps, int can_split, long long partial_capacity, int32_t *out);

// shared-memory-resident family (5 <= n <= 13): pqc_smem.cu
int smem_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv,
                 float gamma, float *d_fitness, double *d_partial, int sm_count, cudaStream_t st);
int smem_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows,
                 float *d_q, cudaStream_t st);
double smem_flops_per_eval(const oqrl_spec &sp);

// register-block form of the shared-memory class (fast path, 5 <= n <= 13): pqc_blk.cu
bool blk_supports(const oqrl_spec &sp);
int blk_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                float *d_fitness, double *d_partial, int sm_count, cudaStream_t st);
int blk_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q,
                cudaStream_t st);
double blk_flops_per_eval(const oqrl_spec &sp);

// HBM-resident family (n >= 14): pqc_hbm.cu
int hbm_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows,
                float *d_q, float *d_state_out, void *workspace, size_t workspace_bytes, int sm_count,
                cudaStream_t st);
size_t hbm_workspace_bytes(const oqrl_spec &sp, int n_circuits);
double hbm_flops_per_eval(const oqrl_spec &sp);
double hbm_sweeps_per_eval(const oqrl_spec &sp);
int hbm_plan_dump(const oqrl_spec &sp, void *buf, size_t cap, size_t *n_bytes, int keep_state);

// introspection shared with the probe library (csrc/probe.cu): api.cu
int work_model(const oqrl_spec *sp, double *flops_per_eval, double *state_sweeps);
int device_sm_count(int *sm_count);

// shared small kernels: api.cu
int launch_trig_table(const float *d_obs, const float *d_next_obs, const int32_t *d_act, const float *d_rew,
                      const float *d_term, int64_t n_rows, int n_feat, float4 *d_trig, float4 *d_meta,
                      cudaStream_t st);
int launch_td_finalize(const double *d_partial, int n_cand, int n_chunks, int n_trans, float *d_fitness,
                       cudaStream_t st);

constexpr int kMaxOut = 32;   // n_out upper bound for the fused TD epilogue
}  // namespace oqrl
