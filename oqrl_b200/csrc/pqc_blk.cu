// Shared-memory-resident PQC kernels, register-block form (5..13 qubits; BASELINE configs 3 and 4).
//
// Same reference code replaced as pqc_smem.cu (/root/reference/src/agent.py:48-61, src/optim.py:64-173).  This is
// the fast path of the shared-memory class; pqc_smem.cu remains as the general fallback (n_out > 8, very deep
// circuits whose coefficient tables do not fit) and as the A/B reference (oqrl_spec.reserved bit 1).
//
// Decomposition: a GROUP of 2^(n-4) threads owns one transition = the two circuits Q(s), Q(s') packed into fp32x2.
// Every thread keeps a block of 16 amplitudes (four "register" qubits, 64 registers) and the thread index supplies
// the other n-4 index bits.  A PASS picks four adjacent index bits [p0, p0+4) as the register bits, loads the
// thread's 16 amplitudes from shared memory, applies the layer's gates of those four wires in registers and writes
// them back: in place for all but the layer's last pass, which instead scatters with the CNOT ring folded into the
// addresses (or the CZ ring folded into signs).  ceil(n/4) shared-memory round trips per layer, no arithmetic
// outside registers.
//
// Every instruction that is not an FFMA2/FMUL2 competes for issue slots and, above ~32 KB of hot code, the FMA pipe
// starves on instruction fetch (scripts/icache_probe.py; the first unrolled-per-layer version of this kernel
// stalled on fetch for 1.9 of every issue cycle).  So ONE pass body serves every pass (block position is a runtime
// value; shared-memory addresses are a Gray-code walk of XORs because the swizzle is GF(2)-linear), the
// re-upload-fused gate coefficients are computed once per transition by the group, layer 0 is folded into the
// generated product state and pruned last-layer gates run as identities instead of branches.
#include "common.cuh"

namespace oqrl {
namespace {

#ifndef OQRL_BLK_WARPS
#define OQRL_BLK_WARPS 16
#endif
constexpr int kBlkMaxOut = 8;

template <int N>
struct BlkCfg {
    static_assert(N >= 5 && N <= 13, "register-block kernels cover 5..13 qubits");
    static constexpr int kGroup = 1 << (N - 4);                       // threads per transition
    static constexpr bool kWarpLocal = kGroup <= 32;                  // group synchronises with __syncwarp
    static constexpr int kThreads = kWarpLocal ? 128 : kGroup;
    static constexpr int kGroups = kThreads / kGroup;                 // transitions in flight per CTA
    static constexpr int kMinCtas = OQRL_BLK_WARPS * 32 / kThreads >= 1 ? OQRL_BLK_WARPS * 32 / kThreads : 1;   // resident warps per SM -> register cap
    static constexpr int kBlocks = (N + 3) / 4;                       // passes per layer
    static constexpr bool kPartial = (N % 4) != 0;                    // last pass overlaps the previous one
    static constexpr int kDim = 1 << N;
};

// dynamic shared memory (state first: 16-byte records, the only bank-sensitive array)
template <int N>
struct BlkSmem {
    using C = BlkCfg<N>;
    float4 *state;       // [groups][2^N]
    float4 *fused;       // [groups][max(L,1)*N][2]  {ar_s, ar_s', ai_s, ai_s'}, {br_s, br_s', bi_s, bi_s'}
    float4 *trig;        // [groups][N]              {c_s, c_s', s_s, s_s'} of the wire's feature
    Su2Scalar *rot;      // [max(L,1)*N]
    unsigned *permcol;   // [N-1][N]  image of index bit b under the CNOT ring of range r
    float *red;          // [warps][2*kBlkMaxOut]
    int fused_stride;
    __device__ BlkSmem(unsigned char *base, int n_layers)
    {
        const int Lc = n_layers > 0 ? n_layers : 1;
        size_t off = 0;
        state = reinterpret_cast<float4 *>(base); off += (size_t)C::kGroups * C::kDim * sizeof(float4);
        fused = reinterpret_cast<float4 *>(base + off); fused_stride = Lc * N * 2; off += (size_t)C::kGroups * fused_stride * sizeof(float4);
        trig = reinterpret_cast<float4 *>(base + off); off += (size_t)C::kGroups * N * sizeof(float4);
        rot = reinterpret_cast<Su2Scalar *>(base + off); off += (size_t)Lc * N * sizeof(Su2Scalar);
        permcol = reinterpret_cast<unsigned *>(base + off); off += (size_t)(N - 1) * N * sizeof(unsigned);
        off = (off + 15) & ~(size_t)15;
        red = reinterpret_cast<float *>(base + off);
    }
    static size_t bytes(int n_layers)
    {
        const int Lc = n_layers > 0 ? n_layers : 1;
        size_t off = (size_t)C::kGroups * C::kDim * sizeof(float4) + (size_t)C::kGroups * Lc * N * 2 * sizeof(float4)
                     + (size_t)C::kGroups * N * sizeof(float4) + (size_t)Lc * N * sizeof(Su2Scalar) + (size_t)(N - 1) * N * sizeof(unsigned);
        off = (off + 15) & ~(size_t)15;
        return off + (size_t)(C::kThreads / 32) * 2 * kBlkMaxOut * sizeof(float);
    }
};

template <int N>
struct BlkCtx {
    const oqrl_spec *sp;
    const Su2Scalar *rot;
    const unsigned *permcol;
    const float4 *trig;     // this group's
    float4 *fused;          // this group's
    unsigned sbase;         // shared-window address of this group's state
    float *red;
    int t;                  // thread inside the group
    unsigned last_mask;
};

template <int N>
__device__ __forceinline__ void group_sync()
{
    if constexpr (BlkCfg<N>::kWarpLocal) __syncwarp();
    else __syncthreads();
}

// Shared-memory slot of basis index k: the low three bits (the 16-byte bank group of the record) are XORed with a
// GF(2)-linear function of index bits 4..6, so slot(a ^ b) = slot(a) ^ slot(b).  The columns come from an offline
// search (scripts/blk_swizzle_search.py): every load / in-place store pattern stays conflict-free (eight lanes on index
// bits 4, 5, 6 for the p0 = 0 block, on bits 0, 1, 2 otherwise) and the relabelling scatter, whose eight lanes land
// on the images of bits 0, 1, 2 under the layer's CNOT ring, conflicts for at most one ring range (2-way).  The first
// choice, k ^ ((k >> 4) & 7), left the scatter 8-way conflicted at range 4 and 2-way at range 6 (n = 8).
template <int N>
__device__ __forceinline__ unsigned blk_slot(unsigned k)
{
    // nibble x of the constant = XOR of the columns selected by x = index bits 4..6
    constexpr unsigned kTable = N == 5 ? 0x20202020u : (N == 6 ? 0x15401540u : 0x14632750u);   // columns (2,0,0) / (4,5,0) / (5,7,3)
    return k ^ ((kTable >> (((k >> 4) & 7u) * 4)) & 7u);
}

// bit that changes between Gray-code steps j-1 and j (j = 1..15)
__device__ __forceinline__ constexpr int gray_bit(int j) { return (j & 1) ? 0 : ((j & 2) ? 1 : ((j & 4) ? 2 : 3)); }

// Group-cooperative: fused coefficient (layers >= 1) or product-state vector (layer 0) of every gate.
template <int N>
__device__ __forceinline__ void blk_build_fused(const BlkCtx<N> &cx)
{
    const oqrl_spec &sp = *cx.sp;
    const int L = sp.n_layers;
    const int total = (L > 0 ? L : 1) * N;
    for (int g = cx.t; g < total; g += BlkCfg<N>::kGroup) {
        const int l = g / N, w = g - l * N;
        const bool emb = embed_feature(w, sp.n_feat, sp.feature_map) >= 0;
        const float4 t = cx.trig[w];
        Su2Scalar r;
        if (L > 0) r = cx.rot[g];
        else { r.ar = 1.f; r.ai = 0.f; r.br = 0.f; r.bi = 0.f; }
        float4 p, q;
        if (l > 0 && l == L - 1 && !((cx.last_mask >> w) & 1u)) {   // outside every read-out's light cone: identity
            p = make_float4(1.f, 1.f, 0.f, 0.f);
            q = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (l == 0) {          // v = [[a, b], [-conj b, conj a]] (c, s)^T
            p = make_float4(fmaf(r.br, t.z, r.ar * t.x), fmaf(r.br, t.w, r.ar * t.y), fmaf(r.bi, t.z, r.ai * t.x), fmaf(r.bi, t.w, r.ai * t.y));
            q = make_float4(fmaf(r.ar, t.z, -r.br * t.x), fmaf(r.ar, t.w, -r.br * t.y), fmaf(-r.ai, t.z, r.bi * t.x), fmaf(-r.ai, t.w, r.bi * t.y));
        } else if (sp.reupload && emb) {   // Rot * RY(x): a' = a c + b s, b' = b c - a s
            p = make_float4(fmaf(r.ar, t.x, r.br * t.z), fmaf(r.ar, t.y, r.br * t.w), fmaf(r.ai, t.x, r.bi * t.z), fmaf(r.ai, t.y, r.bi * t.w));
            q = make_float4(fmaf(r.br, t.x, -r.ar * t.z), fmaf(r.br, t.y, -r.ar * t.w), fmaf(r.bi, t.x, -r.ai * t.z), fmaf(r.bi, t.y, -r.ai * t.w));
        } else {
            p = make_float4(r.ar, r.ar, r.ai, r.ai);
            q = make_float4(r.br, r.br, r.bi, r.bi);
        }
        cx.fused[2 * g] = p;
        cx.fused[2 * g + 1] = q;
    }
}

// one gate (coef -> its two float4 table entries) on register bit Q of the 16-amplitude block
template <int Q>
__device__ __forceinline__ void blk_gate(const float4 *coef, u64 (&re)[16], u64 (&im)[16])
{
    const float4 a = coef[0], b = coef[1];
    Su2Coef g;
    g.ar = pack2(a.x, a.y); g.ai = pack2(a.z, a.w); g.nai = fneg2(g.ai);
    g.br = pack2(b.x, b.y); g.nbr = fneg2(g.br);
    g.bi = pack2(b.z, b.w); g.nbi = fneg2(g.bi);
#pragma unroll
    for (int c = 0; c < 16; ++c)
        if (((c >> Q) & 1) == 0) su2_pair(g, re[c], im[c], re[c | (1 << Q)], im[c | (1 << Q)]);
}

// Whole circuit for the group's transition.  z[i] (packed s / s') valid in thread 0 of the group afterwards.
template <int N>
__device__ __forceinline__ void blk_circuit(const BlkCtx<N> &cx, u64 (&z)[kBlkMaxOut])
{
    using C = BlkCfg<N>;
    const oqrl_spec &sp = *cx.sp;
    const int L = sp.n_layers, t = cx.t;
    const bool sel = sp.entangler == OQRL_ENT_SEL_CNOT;
    const unsigned sbase = cx.sbase;
    u64 re[16], im[16];

    // ---- generated state, block p0 = 0: k = t << 4 | c (thread bit j = index bit 4 + j = wire N-5-j) ----
    {
        u64 pr = pack2(1.f, 1.f), pi = 0ull;
#pragma unroll
        for (int j = 0; j < N - 4; ++j) {
            const int wire = N - 5 - j;
            const float4 v = cx.fused[2 * wire + ((t >> j) & 1)];
            const u64 vr = pack2(v.x, v.y), vi = pack2(v.z, v.w);
            const u64 nr = ffma2(fneg2(pi), vi, fmul2(pr, vr));
            pi = ffma2(pi, vr, fmul2(pr, vi));
            pr = nr;
        }
        re[0] = pr; im[0] = pi;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int wire = N - 1 - q;
            const float4 v0 = cx.fused[2 * wire], v1 = cx.fused[2 * wire + 1];
            const u64 v0r = pack2(v0.x, v0.y), v0i = pack2(v0.z, v0.w), nv0i = fneg2(v0i);
            const u64 v1r = pack2(v1.x, v1.y), v1i = pack2(v1.z, v1.w), nv1i = fneg2(v1i);
#pragma unroll
            for (int c = (1 << q) - 1; c >= 0; --c) {
                const u64 xr = re[c], xi = im[c];
                re[c] = ffma2(xi, nv0i, fmul2(xr, v0r));
                im[c] = ffma2(xi, v0r, fmul2(xr, v0i));
                re[c | (1 << q)] = ffma2(xi, nv1i, fmul2(xr, v1r));
                im[c | (1 << q)] = ffma2(xi, v1r, fmul2(xr, v1i));
            }
        }
    }

    // thread part of the basis index for the block at p0: thread bits below p0 stay, the rest move up by four
    auto thread_index = [&](int p0) { return (unsigned)(((t >> p0) << (p0 + 4)) | (t & ((1 << p0) - 1))); };
    // byte offsets of the block's 16 records: Gray-code walk from the thread's own offset
    auto load_block = [&](int p0) {
        unsigned o = blk_slot<N>(thread_index(p0)) << 4, se[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) se[q] = blk_slot<N>(1u << (p0 + q)) << 4;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = j ^ (j >> 1);
            if (j > 0) o ^= se[gray_bit(j)];
            ld128(sbase + o, re[c], im[c]);
        }
    };
    auto store_block = [&](int p0) {
        unsigned o = blk_slot<N>(thread_index(p0)) << 4, se[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) se[q] = blk_slot<N>(1u << (p0 + q)) << 4;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = j ^ (j >> 1);
            if (j > 0) o ^= se[gray_bit(j)];
            st128(sbase + o, re[c], im[c]);
        }
    };
    // Images under the layer's CNOT ring M (identity for CZ / no layer) of the block's register bits (mc) and of
    // the thread's part of the index (mu).
    auto relabel_columns = [&](int layer, int p0, bool relabel, unsigned (&mc)[4], unsigned &mu) {
        const unsigned *pc = cx.permcol + (relabel ? (sel_range(N, layer) - 1) * N : 0);
        mu = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) mc[q] = relabel ? pc[p0 + q] : (1u << (p0 + q));
#pragma unroll
        for (int j = 0; j < N - 4; ++j) {
            const int b = j < p0 ? j : j + 4;
            if ((t >> j) & 1) mu ^= relabel ? pc[b] : (1u << b);
        }
    };
    // Layer's entangler folded into the write-back.  CNOT ring: element k goes to the slot of M k (one XOR per
    // element, the slot function being linear).  CZ ring: same slot, sign (-1)^{#cyclically adjacent 11 pairs}.
    auto scatter_entangled = [&](int layer, int p0) {
        if (sel) {
            unsigned mc[4], mu, sm4[4];
            relabel_columns(layer, p0, true, mc, mu);
#pragma unroll
            for (int q = 0; q < 4; ++q) sm4[q] = blk_slot<N>(mc[q]) << 4;
            unsigned o = blk_slot<N>(mu) << 4;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = j ^ (j >> 1);
                if (j > 0) o ^= sm4[gray_bit(j)];
                st128(sbase + o, re[c], im[c]);
            }
        } else {
            const unsigned tk = thread_index(p0);
            unsigned o = blk_slot<N>(tk) << 4, se[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) se[q] = blk_slot<N>(1u << (p0 + q)) << 4;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = j ^ (j >> 1);
                if (j > 0) o ^= se[gray_bit(j)];
                const unsigned kk = tk | ((unsigned)c << p0);
                const unsigned rotl = ((kk << 1) | (kk >> (N - 1))) & (unsigned)(C::kDim - 1);
                const u64 mk = (__popc(kk & rotl) & 1) ? 0x8000000080000000ull : 0ull;
                st128(sbase + o, re[c] ^ mk, im[c] ^ mk);
            }
        }
    };
    // <Z_i> after the last entangler: signed sum of |amp|^2 with sign = bit (N-1-i) of M k.  The sign is linear in
    // the register index, so the 16-term sum is a 4-level butterfly of FMAs with +-1 multipliers.
    auto readout = [&](int layer, int p0) {
        unsigned mc[4], mu;
        relabel_columns(layer, p0, sel && layer >= 0, mc, mu);
        u64 p[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) p[c] = ffma2(im[c], im[c], fmul2(re[c], re[c]));
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < kBlkMaxOut; ++i) {
            z[i] = 0ull;
            if (i < sp.n_out) {
                u64 s[8];
                float sg[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) sg[q] = ((mc[q] >> (N - 1 - i)) & 1u) ? -1.f : 1.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) s[c] = ffma2(p[2 * c + 1], bcast2(sg[0]), p[2 * c]);
#pragma unroll
                for (int c = 0; c < 4; ++c) s[c] = ffma2(s[2 * c + 1], bcast2(sg[1]), s[2 * c]);
#pragma unroll
                for (int c = 0; c < 2; ++c) s[c] = ffma2(s[2 * c + 1], bcast2(sg[2]), s[2 * c]);
                s[0] = ffma2(s[1], bcast2(sg[3]), s[0]);
                float lo, hi;
                unpack2(s[0], lo, hi);
                if ((mu >> (N - 1 - i)) & 1u) { lo = -lo; hi = -hi; }
#pragma unroll
                for (int off = (C::kGroup < 32 ? C::kGroup : 32) >> 1; off > 0; off >>= 1) {
                    lo += __shfl_xor_sync(0xffffffffu, lo, off);
                    hi += __shfl_xor_sync(0xffffffffu, hi, off);
                }
                if constexpr (!C::kWarpLocal) {
                    if (lane == 0) { cx.red[(warp * kBlkMaxOut + i) * 2] = lo; cx.red[(warp * kBlkMaxOut + i) * 2 + 1] = hi; }
                } else {
                    z[i] = pack2(lo, hi);
                }
            }
        }
        if constexpr (!C::kWarpLocal) {
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int i = 0; i < sp.n_out; ++i) {
                    float lo = 0.f, hi = 0.f;
                    for (int w = 0; w < C::kThreads / 32; ++w) { lo += cx.red[(w * kBlkMaxOut + i) * 2]; hi += cx.red[(w * kBlkMaxOut + i) * 2 + 1]; }
                    cx.red[i * 2] = lo; cx.red[i * 2 + 1] = hi;     // thread 0 owns warp 0's row: no race
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kBlkMaxOut; ++i)
                if (i < sp.n_out) z[i] = pack2(cx.red[i * 2], cx.red[i * 2 + 1]);
            __syncthreads();   // red is rewritten by the next transition
        }
    };

    if (L <= 1) {                       // no further layer: read the generated state out (through layer 0's entangler)
        readout(L == 1 ? 0 : -1, 0);
        return;
    }
    scatter_entangled(0, 0);
    group_sync<N>();
    // one pass = load the block at p0, the layer's gates on its four register bits, store (in place, or -- the layer's
    // last pass -- scattered through the entangler)
    auto pass = [&](int l, int bi, bool last_of_circuit) {
        const int p0 = (C::kPartial && bi == C::kBlocks - 1) ? N - 4 : 4 * bi;
        const int start_q = C::kPartial ? 4 * bi - p0 : 0;      // register bits below start_q were done by the previous pass
        load_block(p0);
        const float4 *coef = cx.fused + 2 * (l * N + N - 1 - p0);   // register bit q <-> index bit p0 + q <-> wire N-1-p0-q
        if (!C::kPartial || start_q <= 0) blk_gate<0>(coef, re, im);
        if (!C::kPartial || start_q <= 1) blk_gate<1>(coef - 2, re, im);
        if (!C::kPartial || start_q <= 2) blk_gate<2>(coef - 4, re, im);
        blk_gate<3>(coef - 6, re, im);
        if (bi != C::kBlocks - 1) {
            store_block(p0);                               // same thread, same slots as load_block: no hazard before it
            group_sync<N>();
        } else if (!last_of_circuit) {
            group_sync<N>();
            scatter_entangled(l, p0);
            group_sync<N>();
        }
    };
    if constexpr (C::kBlocks == 2 && !C::kPartial) {
        // two passes per layer (8 qubits): the layer is one straight-line body -- no branch between the two kinds of
        // store, and the registers only have to be back in place once per layer
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
            pass(l, 0, false);
            pass(l, 1, l == L - 1);
        }
    } else {
        const int n_pass = (L - 1) * C::kBlocks;
        int l = 1, bi = 0;
#pragma unroll 1
        for (int j = 0; j < n_pass; ++j) {
            pass(l, bi, j == n_pass - 1);
            if (bi != C::kBlocks - 1) ++bi;
            else { bi = 0; ++l; }
        }
    }
    readout(L - 1, (C::kPartial) ? N - 4 : 4 * (C::kBlocks - 1));
}

template <int N>
__device__ __forceinline__ void blk_prologue(const oqrl_spec &sp, const float *__restrict__ params, BlkSmem<N> &sm)
{
    const int n_gates = sp.n_layers * N;
    for (int g = threadIdx.x; g < n_gates; g += blockDim.x) {
        Su2Scalar c;
        rot_su2(params[3 * g], params[3 * g + 1], params[3 * g + 2], c.ar, c.ai, c.br, c.bi);
        sm.rot[g] = c;
    }
    for (int e = threadIdx.x; e < (N - 1) * N; e += blockDim.x) {
        const int r = e / N + 1, b = e % N;
        sm.permcol[e] = sel_perm(1u << b, N, r);
    }
}

template <int N>
__device__ __forceinline__ BlkCtx<N> blk_context(const oqrl_spec &sp, const BlkSmem<N> &sm)
{
    using C = BlkCfg<N>;
    const int group = threadIdx.x / C::kGroup, t = threadIdx.x % C::kGroup;
    const unsigned last_mask = (sp.reserved & 1) ? ~0u : readout_wire_mask(N, sp.n_layers, sp.n_out, sp.entangler);
    BlkCtx<N> cx;
    cx.sp = &sp;
    cx.rot = sm.rot;
    cx.permcol = sm.permcol;
    cx.trig = sm.trig + group * N;
    cx.fused = sm.fused + (size_t)group * sm.fused_stride;
    cx.sbase = (unsigned)__cvta_generic_to_shared(sm.state + (size_t)group * C::kDim);
    cx.red = sm.red;
    cx.t = t;
    cx.last_mask = last_mask;
    return cx;
}

template <int N>
__global__ void __launch_bounds__(BlkCfg<N>::kThreads, BlkCfg<N>::kMinCtas)
blk_fitness_kernel(oqrl_spec sp, const float *__restrict__ params, TransitionView tv, float gamma, int chunk_len,
                   float *__restrict__ fitness, double *__restrict__ partial)
{
    using C = BlkCfg<N>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double warp_part[C::kThreads / 32];
    BlkSmem<N> sm(smem_raw, sp.n_layers);
    const int cand = blockIdx.x;
    blk_prologue<N>(sp, params + (size_t)cand * sp.n_layers * N * 3, sm);
    __syncthreads();
    const BlkCtx<N> cx = blk_context<N>(sp, sm);
    const int group = threadIdx.x / C::kGroup, t = cx.t;
    float4 *my_trig = sm.trig + group * N;

    const int t0 = blockIdx.y * chunk_len;
    const int t1 = min(tv.n_trans, t0 + chunk_len);
    double acc = 0.0;   // fp64: the sum then depends on how transitions are chunked only at the 1e-16 level
    bool bad_idx = false;
    for (int tb = t0; tb < t1; tb += C::kGroups) {            // trip count is CTA-uniform
        const int tr = tb + group;
        const bool valid = tr < t1;
        const int row = tv.idx ? checked_row(tv.idx[valid ? tr : t0], tv.n_rows, bad_idx) : (valid ? tr : t0);
        for (int w = t; w < N; w += C::kGroup) {
            const int f = embed_feature(w, sp.n_feat, sp.feature_map);
            my_trig[w] = f >= 0 ? __ldg(tv.trig + (size_t)row * sp.n_feat + f) : make_float4(1.f, 1.f, 0.f, 0.f);
        }
        group_sync<N>();
        blk_build_fused<N>(cx);
        group_sync<N>();
        u64 z[kBlkMaxOut];
        blk_circuit<N>(cx, z);
        if (t == 0 && valid) {
            const float4 meta = tv.meta[row];
            const int act = __float_as_int(meta.z);
            float qa = 0.f, mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < kBlkMaxOut; ++i) {
                if (i < sp.n_out) {
                    float lo, hi;
                    unpack2(z[i], lo, hi);
                    if (i == act) qa = lo;
                    mx = fmaxf(mx, hi);
                }
            }
            const float target = meta.x + (1.f - meta.y) * gamma * mx;   // optim.py:163
            const float d = qa - target;
            acc += (double)(d * d);
        }
        group_sync<N>();   // trig / fused / state are reused by the next transition
    }
    double v = bad_idx ? (double)NAN : acc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < C::kThreads / 32; ++w) tot += warp_part[w];
        if (gridDim.y == 1) fitness[cand] = (float)(-tot / (double)tv.n_trans);
        else partial[(size_t)cand * gridDim.y + blockIdx.y] = tot;
    }
}

template <int N>
__global__ void __launch_bounds__(BlkCfg<N>::kThreads, BlkCfg<N>::kMinCtas)
blk_qvalues_kernel(oqrl_spec sp, const float *__restrict__ params, const float *__restrict__ x, int n_rows, float *__restrict__ q)
{
    using C = BlkCfg<N>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    BlkSmem<N> sm(smem_raw, sp.n_layers);
    const int cand = blockIdx.x;
    blk_prologue<N>(sp, params + (size_t)cand * sp.n_layers * N * 3, sm);
    __syncthreads();
    const BlkCtx<N> cx = blk_context<N>(sp, sm);
    const int group = threadIdx.x / C::kGroup, t = cx.t;
    float4 *my_trig = sm.trig + group * N;
    const int n_pairs = (n_rows + 1) >> 1;
    for (int jb = blockIdx.y * C::kGroups; jb < n_pairs; jb += gridDim.y * C::kGroups) {
        const int j = jb + group;
        const bool valid = j < n_pairs;
        const int b0 = valid ? 2 * j : 0, b1 = valid ? min(2 * j + 1, n_rows - 1) : 0;
        for (int w = t; w < N; w += C::kGroup) {
            const int f = embed_feature(w, sp.n_feat, sp.feature_map);
            float4 tg = make_float4(1.f, 1.f, 0.f, 0.f);
            if (f >= 0) {
                sincosf(0.5f * x[(size_t)b0 * sp.n_feat + f], &tg.z, &tg.x);
                sincosf(0.5f * x[(size_t)b1 * sp.n_feat + f], &tg.w, &tg.y);
            }
            my_trig[w] = tg;
        }
        group_sync<N>();
        blk_build_fused<N>(cx);
        group_sync<N>();
        u64 z[kBlkMaxOut];
        blk_circuit<N>(cx, z);
        if (t == 0 && valid) {
#pragma unroll
            for (int i = 0; i < kBlkMaxOut; ++i) {
                if (i < sp.n_out) {
                    float lo, hi;
                    unpack2(z[i], lo, hi);
                    q[((size_t)cand * n_rows + b0) * sp.n_out + i] = lo;
                    if (b1 != b0) q[((size_t)cand * n_rows + b1) * sp.n_out + i] = hi;
                }
            }
        }
        group_sync<N>();
    }
}

template <int N>
int launch_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                   float *d_fitness, double *d_partial, int sm_count, cudaStream_t st)
{
    using C = BlkCfg<N>;
    const size_t smem = BlkSmem<N>::bytes(sp.n_layers);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(blk_fitness_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int T = tv.n_trans;
    int chunks = 1;
    if (d_partial) {
        const int want = (4 * sm_count + n_cand - 1) / n_cand;
        const int max_chunks = (T + C::kGroups - 1) / C::kGroups;
        chunks = want < 1 ? 1 : (want > max_chunks ? max_chunks : want);
        if (chunks > 4 * sm_count) chunks = 4 * sm_count;
    }
    int chunk_len = (T + chunks - 1) / chunks;
    chunk_len = ((chunk_len + C::kGroups - 1) / C::kGroups) * C::kGroups;
    chunks = (T + chunk_len - 1) / chunk_len;
    dim3 grid(n_cand, chunks);
    blk_fitness_kernel<N><<<grid, C::kThreads, smem, st>>>(sp, d_params, tv, gamma, chunk_len, d_fitness, d_partial);
    OQRL_LAUNCH_CHECK("blk_fitness_kernel");
    if (chunks > 1) return launch_td_finalize(d_partial, n_cand, chunks, T, d_fitness, st);
    return OQRL_OK;
}

template <int N>
int launch_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q, cudaStream_t st)
{
    using C = BlkCfg<N>;
    const size_t smem = BlkSmem<N>::bytes(sp.n_layers);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(blk_qvalues_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_pairs = (n_rows + 1) / 2;
    int gy = (n_pairs + C::kGroups - 1) / C::kGroups;
    if (gy > 2048) gy = 2048;
    dim3 grid(n_cand, gy);
    blk_qvalues_kernel<N><<<grid, C::kThreads, smem, st>>>(sp, d_params, d_x, n_rows, d_q);
    OQRL_LAUNCH_CHECK("blk_qvalues_kernel");
    return OQRL_OK;
}

template <int N>
size_t smem_bytes(int n_layers) { return BlkSmem<N>::bytes(n_layers); }

#define OQRL_BLK_DISPATCH(FN, n, ...)                                                            \
    switch (n) {                                                                                 \
    case 5: return FN<5>(__VA_ARGS__);                                                           \
    case 6: return FN<6>(__VA_ARGS__);                                                           \
    case 7: return FN<7>(__VA_ARGS__);                                                           \
    case 8: return FN<8>(__VA_ARGS__);                                                           \
    case 9: return FN<9>(__VA_ARGS__);                                                           \
    case 10: return FN<10>(__VA_ARGS__);                                                         \
    case 11: return FN<11>(__VA_ARGS__);                                                         \
    case 12: return FN<12>(__VA_ARGS__);                                                         \
    case 13: return FN<13>(__VA_ARGS__);                                                         \
    default: break;                                                                              \
    }

size_t blk_smem_bytes(int n, int n_layers)
{
    OQRL_BLK_DISPATCH(smem_bytes, n, n_layers);
    return (size_t)-1;
}

}  // namespace

// The register-block kernels take the circuit when its tables fit next to the state (the per-transition fused
// coefficient tables grow with L and with the number of transitions in flight per CTA).
bool blk_supports(const oqrl_spec &sp)
{
    if (sp.n_qubits < 5 || sp.n_qubits > 13 || sp.n_out > kBlkMaxOut) return false;
    const size_t cap = sp.n_qubits >= 12 ? 200 * 1024 : 72 * 1024;   // keep >= 3 CTAs of 128 threads per SM
    return blk_smem_bytes(sp.n_qubits, sp.n_layers) <= cap;
}

int blk_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                float *d_fitness, double *d_partial, int sm_count, cudaStream_t st)
{
    OQRL_BLK_DISPATCH(launch_fitness, sp.n_qubits, sp, d_params, n_cand, tv, gamma, d_fitness, d_partial, sm_count, st);
    return OQRL_ERR_INVALID;
}

int blk_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q, cudaStream_t st)
{
    OQRL_BLK_DISPATCH(launch_qvalues, sp.n_qubits, sp, d_params, n_cand, d_x, n_rows, d_q, st);
    return OQRL_ERR_INVALID;
}

// Executed fp32 flop per circuit evaluation.
double blk_flops_per_eval(const oqrl_spec &sp)
{
    const int n = sp.n_qubits, L = sp.n_layers;
    const double dim = (double)(1u << n);
    int n_emb = 0;
    for (int w = 0; w < n; ++w) n_emb += embed_feature(w, sp.n_feat, sp.feature_map) >= 0;
    double f = 6.0 * (dim / 16.0) * ((n - 4) + 30.0);  // generated state: per thread n-4 + 30 complex multiplies
    f += 18.0 * n;                                     // per-wire product vectors (layer 0 folded in)
    if (L > 1) {
        f += 14.0 * dim * n * (L - 1);               // SU(2) gates, 28 flop per amplitude pair (re-uploaded RY fused in;
                                                       // last-layer gates outside the light cone run as identities)
        if (sp.reupload) f += 12.0 * n_emb * (L - 1);  // one Rot * RY product per gate per transition
    }
    f += 3.0 * dim + (double)sp.n_out * dim;           // |amp|^2 and signed sums
    return f;
}

}  // namespace oqrl
