// Shared-memory-resident PQC kernels for 5 <= n <= 13 qubits (BASELINE config 3 and its neighbours).
//
// Replaces the same reference code as pqc_reg.cu (/root/reference/src/agent.py:48-61 and
// /root/reference/src/optim.py:64-173) for circuits whose state no longer fits one thread.
//
// Decomposition
//   * a CTA belongs to one candidate (gate table CTA-uniform in shared memory);
//   * a GROUP of 2^(n-4) threads owns one transition: the 2^n-amplitude state of BOTH circuits
//     (Q(s) low half / Q(s') high half of packed fp32x2) lives in shared memory as 16-byte
//     records {re_s, re_s', im_s, im_s'} -- one LDS.128 / STS.128 per amplitude;
//   * a PASS register-blocks four index bits: every thread pulls a 16-amplitude block
//     (64 registers), applies all of the layer's gates that act on those four wires
//     (8 amplitude pairs x 16 FFMA2/FMUL2 per gate), and writes the block back, so a layer
//     costs ceil(n/4) state sweeps through shared memory instead of n;
//   * the CNOT ring is a GF(2)-linear relabelling of basis states: it is folded into the store
//     addresses of the layer's last pass (0 flop); the CZ ring is a sign folded into the same store;
//   * data re-uploading fuses RY(x_w) into the following Rot (one SU(2) product per gate per
//     thread instead of a second sweep);
//   * the first pass generates the embedded product state in registers, the last pass reduces
//     |amp|^2 into <Z_i> without storing the final state.
// Records are XOR-swizzled (slot = j ^ ((j >> 4) & 7)) so the 8 lanes of an LDS.128 phase hit
// 8 distinct 16-byte bank groups for every block position.
#include "common.cuh"

namespace oqrl {
namespace {

constexpr int kSmemMaxGates = 1024;   // L * n bound for this family (64 KiB of coefficients)
constexpr int kSmemMaxOut = 8;        // <Z_i> outputs kept in registers

template <int N>
struct Cfg {
    static_assert(N >= 5 && N <= 13, "shared-memory family covers 5..13 qubits");
    static constexpr int kGroup = 1 << (N - 4);                     // threads per transition
    static constexpr int kThreads = kGroup > 32 ? kGroup : 128;   // multi-warp groups own the CTA (group_sync = __syncthreads)
    static constexpr int kGroupsPerCta = kThreads / kGroup;
    static constexpr int kPasses = (N + 3) / 4;
    static constexpr int kDim = 1 << N;
};

__device__ __forceinline__ int swz(int j) { return j ^ ((j >> 4) & 7); }

struct GateRaw { float ar, ai, br, bi; };   // SU(2): [[a, b], [-conj(b), conj(a)]]

// Dynamic shared memory carve-up (per CTA):
//   GateRaw gates[n_gates]            Rot gates of the candidate
//   u32     perm_col[(N-1)][N]        images of the unit vectors under each layer's CNOT ring
//   float4  trig[kGroupsPerCta][N]    packed half-angle pairs {c_s, c_s', s_s, s_s'} of the wire's feature
//   float4  state[kGroupsPerCta][2^N] amplitudes
template <int N>
struct SmemLayout {
    GateRaw *gates;
    unsigned *perm_col;
    float4 *trig;
    float4 *state;
    __device__ SmemLayout(unsigned char *base, int n_gates)
    {
        gates = reinterpret_cast<GateRaw *>(base);
        size_t off = ((size_t)n_gates * sizeof(GateRaw) + 15) & ~(size_t)15;
        perm_col = reinterpret_cast<unsigned *>(base + off);
        off += (((size_t)(N - 1) * N * sizeof(unsigned)) + 15) & ~(size_t)15;
        trig = reinterpret_cast<float4 *>(base + off);
        off += (size_t)Cfg<N>::kGroupsPerCta * N * sizeof(float4);
        state = reinterpret_cast<float4 *>(base + off);
    }
    static size_t bytes(int n_gates)
    {
        size_t off = ((size_t)n_gates * sizeof(GateRaw) + 15) & ~(size_t)15;
        off += (((size_t)(N - 1) * N * sizeof(unsigned)) + 15) & ~(size_t)15;
        off += (size_t)Cfg<N>::kGroupsPerCta * N * sizeof(float4);
        off += (size_t)Cfg<N>::kGroupsPerCta * sizeof(float4) * Cfg<N>::kDim;
        return off;
    }
};

template <int N>
__device__ __forceinline__ void group_sync()
{
    if constexpr (Cfg<N>::kGroup <= 32) __syncwarp();
    else __syncthreads();   // kGroup == blockDim.x in that case
}

// One pass over block positions [lo, lo+4): gates act on positions >= gate_lo.
// FIRST: the block is generated (embedded product state) instead of loaded.
// LAST_OF_LAYER: entangler folded into the store; FINAL: probabilities reduced instead of stored.
template <int N>
struct PassCtx {
    const oqrl_spec *sp;
    const GateRaw *gates;      // [L][N]
    const unsigned *perm_col;  // [(N-1)][N]
    const float4 *trig;        // this group's [N]
    float4 *state;             // this group's [2^N]
    int u;                     // thread index inside the group
};

template <int N>
__device__ __forceinline__ void load_coef(const PassCtx<N> &cx, int layer, int wire, bool fuse_ry, Su2Coef &g)
{
    const GateRaw r = cx.gates[layer * N + wire];
    if (fuse_ry) {
        // U = Rot * RY(x):  a' = a c + b s,  b' = b c - a s   (per circuit: lo = s, hi = s')
        const float4 t = cx.trig[wire];
        const float c0 = t.x, c1 = t.y, s0 = t.z, s1 = t.w;
        const float ar0 = fmaf(r.ar, c0, r.br * s0), ar1 = fmaf(r.ar, c1, r.br * s1);
        const float ai0 = fmaf(r.ai, c0, r.bi * s0), ai1 = fmaf(r.ai, c1, r.bi * s1);
        const float br0 = fmaf(r.br, c0, -r.ar * s0), br1 = fmaf(r.br, c1, -r.ar * s1);
        const float bi0 = fmaf(r.bi, c0, -r.ai * s0), bi1 = fmaf(r.bi, c1, -r.ai * s1);
        g.ar = pack2(ar0, ar1); g.ai = pack2(ai0, ai1); g.nai = pack2(-ai0, -ai1);
        g.br = pack2(br0, br1); g.nbr = pack2(-br0, -br1);
        g.bi = pack2(bi0, bi1); g.nbi = pack2(-bi0, -bi1);
    } else {
        g.ar = pack2(r.ar, r.ar); g.ai = pack2(r.ai, r.ai); g.nai = pack2(-r.ai, -r.ai);
        g.br = pack2(r.br, r.br); g.nbr = pack2(-r.br, -r.br);
        g.bi = pack2(r.bi, r.bi); g.nbi = pack2(-r.bi, -r.bi);
    }
}

// Apply the gate of block-local bit q (compile-time) to the 16-amplitude block.
template <int Q>
__device__ __forceinline__ void block_gate(const Su2Coef &g, u64 (&re)[16], u64 (&im)[16])
{
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        if (((c >> Q) & 1) == 0) su2_pair(g, re[c], im[c], re[c | (1 << Q)], im[c | (1 << Q)]);
    }
}

template <int N, int NOUT_CAP>
__device__ __forceinline__ void run_pass(const PassCtx<N> &cx, int layer, int pass, bool first, bool last_of_layer,
                                         bool final_pass, u64 (&zacc)[NOUT_CAP])
{
    const oqrl_spec &sp = *cx.sp;
    const int lo = min(4 * pass, N - 4);        // lowest block bit position
    const int gate_lo = 4 * pass;               // positions below this were handled by an earlier pass
    const int u = cx.u;
    const int jbase = ((u >> lo) << (lo + 4)) | (u & ((1 << lo) - 1));
    u64 re[16], im[16];

    if (first) {
        // embedded product state: amp(k) = prod_w (bit_w(k) ? sin : cos) over embedded wires, 0/1 elsewhere
        u64 outer = pack2(1.f, 1.f);
#pragma unroll
        for (int pos = 0; pos < N; ++pos) {
            if (pos < lo || pos >= lo + 4) {
                const int wire = N - 1 - pos;
                const bool bit = (jbase >> pos) & 1;
                const bool emb = embed_feature(wire, sp.n_feat, sp.feature_map) >= 0;
                const float4 t = cx.trig[wire];
                const u64 f = emb ? (bit ? pack2(t.z, t.w) : pack2(t.x, t.y)) : (bit ? 0ull : pack2(1.f, 1.f));
                outer = fmul2(outer, f);
            }
        }
        re[0] = outer;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int wire = N - 1 - (lo + q);
            const bool emb = embed_feature(wire, sp.n_feat, sp.feature_map) >= 0;
            const float4 t = cx.trig[wire];
            const u64 cw = emb ? pack2(t.x, t.y) : pack2(1.f, 1.f);
            const u64 sw = emb ? pack2(t.z, t.w) : 0ull;
#pragma unroll
            for (int c = (1 << q) - 1; c >= 0; --c) {
                const u64 v = re[c];
                re[c] = fmul2(v, cw);
                re[c | (1 << q)] = fmul2(v, sw);
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) im[c] = 0ull;
    } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 v = cx.state[swz(jbase | (c << lo))];
            re[c] = pack2(v.x, v.y);
            im[c] = pack2(v.z, v.w);
        }
    }

    // gates of this layer on the block's wires (positions >= gate_lo only); layer < 0 = circuit without layers
    const bool has_layer = layer >= 0;
    const bool fuse = sp.reupload && layer > 0;
#define OQRL_BLOCK_GATE(Q)                                                                     \
    if (has_layer && lo + Q >= gate_lo) {                                                                   \
        const int wire = N - 1 - (lo + Q);                                                     \
        Su2Coef g;                                                                             \
        load_coef<N>(cx, layer, wire, fuse && embed_feature(wire, sp.n_feat, sp.feature_map) >= 0, g); \
        block_gate<Q>(g, re, im);                                                              \
    }
    OQRL_BLOCK_GATE(0)
    OQRL_BLOCK_GATE(1)
    OQRL_BLOCK_GATE(2)
    OQRL_BLOCK_GATE(3)
#undef OQRL_BLOCK_GATE

    if (!last_of_layer) {
        // in place: every thread rewrites exactly the records it read
        if (first) group_sync<N>();   // nothing was read; keep the barrier count uniform
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float4 v;
            unpack2(re[c], v.x, v.y);
            unpack2(im[c], v.z, v.w);
            cx.state[swz(jbase | (c << lo))] = v;
        }
        group_sync<N>();
        return;
    }

    // entangler folded into the destination index / sign
    int dst_base = jbase;
    int col[4] = {1 << lo, 2 << lo, 4 << lo, 8 << lo};
    const bool sel = has_layer && sp.entangler == OQRL_ENT_SEL_CNOT;
    const bool cz = has_layer && sp.entangler == OQRL_ENT_CZ_RING;
    if (sel) {
        const unsigned *pc = cx.perm_col + (sel_range(N, layer) - 1) * N;
        dst_base = 0;
#pragma unroll
        for (int pos = 0; pos < N; ++pos)
            if ((jbase >> pos) & 1) dst_base ^= pc[pos];
#pragma unroll
        for (int q = 0; q < 4; ++q) col[q] = pc[lo + q];
    }
    if (!final_pass) {
        group_sync<N>();   // all loads of this pass are done before any scattered store lands
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            int dst = dst_base;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((c >> q) & 1) dst ^= col[q];
            u64 r = re[c], i = im[c];
            if (cz) {
                const unsigned k = (unsigned)dst;
                const unsigned rot = ((k << 1) | (k >> (N - 1))) & ((1u << N) - 1);   // bit i meets bit i+1 (ring)
                if (__popc(k & rot) & 1) { r = fneg2(r); i = fneg2(i); }
            }
            float4 v;
            unpack2(r, v.x, v.y);
            unpack2(i, v.z, v.w);
            cx.state[swz(dst)] = v;
        }
        group_sync<N>();
        return;
    }
    // final pass of the final layer: <Z_i> partial sums, sign from the relabelled index
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        int dst = dst_base;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if ((c >> q) & 1) dst ^= col[q];
        const u64 p = ffma2(im[c], im[c], fmul2(re[c], re[c]));
#pragma unroll
        for (int i = 0; i < NOUT_CAP; ++i) {
            if (i < sp.n_out) {
                const bool neg = (dst >> (N - 1 - i)) & 1;
                zacc[i] = neg ? fsub2(zacc[i], p) : fadd2(zacc[i], p);
            }
        }
    }
}

// Whole circuit for the group's transition; returns <Z_i> (packed s / s') valid in every thread of the group.
template <int N>
__device__ __forceinline__ void run_circuit(const PassCtx<N> &cx, u64 (&z)[kSmemMaxOut], float *red_scratch)
{
    const oqrl_spec &sp = *cx.sp;
    constexpr int kPasses = Cfg<N>::kPasses;
#pragma unroll
    for (int i = 0; i < kSmemMaxOut; ++i) z[i] = 0ull;
    const int L = sp.n_layers;
    // L == 0: the embedded product state is read out directly (layer index -1 = no gates, no entangler)
    for (int l = 0; l < max(L, 1); ++l) {
        for (int p = 0; p < kPasses; ++p) {
            const bool first = (l == 0 && p == 0);
            const bool last = (p == kPasses - 1);
            run_pass<N, kSmemMaxOut>(cx, L == 0 ? -1 : l, p, first, last, last && l == max(L, 1) - 1, z);
        }
    }
    // reduce the partial sums over the group
    if constexpr (Cfg<N>::kGroup <= 32) {
#pragma unroll
        for (int i = 0; i < kSmemMaxOut; ++i) {
            if (i < sp.n_out) {
                float lo, hi;
                unpack2(z[i], lo, hi);
#pragma unroll
                for (int off = Cfg<N>::kGroup / 2; off > 0; off >>= 1) {
                    lo += __shfl_xor_sync(0xffffffffu, lo, off);
                    hi += __shfl_xor_sync(0xffffffffu, hi, off);
                }
                z[i] = pack2(lo, hi);
            }
        }
    } else {
        // group == CTA: warp shuffle, then a fixed-order sum over warps through shared memory
        constexpr int kWarps = Cfg<N>::kGroup / 32;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < kSmemMaxOut; ++i) {
            if (i < sp.n_out) {
                float lo, hi;
                unpack2(z[i], lo, hi);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    lo += __shfl_xor_sync(0xffffffffu, lo, off);
                    hi += __shfl_xor_sync(0xffffffffu, hi, off);
                }
                if (lane == 0) { red_scratch[(i * kWarps + warp) * 2] = lo; red_scratch[(i * kWarps + warp) * 2 + 1] = hi; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kSmemMaxOut; ++i) {
            if (i < sp.n_out) {
                float lo = 0.f, hi = 0.f;
                for (int w = 0; w < kWarps; ++w) { lo += red_scratch[(i * kWarps + w) * 2]; hi += red_scratch[(i * kWarps + w) * 2 + 1]; }
                z[i] = pack2(lo, hi);
            }
        }
        __syncthreads();
    }
}

template <int N>
__device__ __forceinline__ void cta_prologue(const oqrl_spec &sp, const float *__restrict__ params, SmemLayout<N> &sm)
{
    const int n_gates = sp.n_layers * N;
    for (int g = threadIdx.x; g < n_gates; g += blockDim.x) {
        GateRaw r;
        rot_su2(params[3 * g], params[3 * g + 1], params[3 * g + 2], r.ar, r.ai, r.br, r.bi);
        sm.gates[g] = r;
    }
    for (int e = threadIdx.x; e < (N - 1) * N; e += blockDim.x) {
        const int r = e / N + 1, pos = e % N;
        sm.perm_col[e] = sel_perm(1u << pos, N, r);
    }
}

__device__ __forceinline__ double block_sum_d(double v, double *warp_part)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) tot += warp_part[w];
    }
    return tot;
}

template <int N>
__global__ void __launch_bounds__(Cfg<N>::kThreads)
smem_fitness_kernel(oqrl_spec sp, const float *__restrict__ params, TransitionView tv, float gamma, int chunk_len,
                    float *__restrict__ fitness, double *__restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double warp_part[32];
    __shared__ float red_scratch[kSmemMaxOut * 32 * 2];
    using C = Cfg<N>;
    SmemLayout<N> sm(smem_raw, sp.n_layers * N);
    const int cand = blockIdx.x;
    cta_prologue<N>(sp, params + (size_t)cand * sp.n_layers * N * 3, sm);
    __syncthreads();

    const int group = threadIdx.x / C::kGroup, u = threadIdx.x % C::kGroup;
    PassCtx<N> cx{&sp, sm.gates, sm.perm_col, sm.trig + group * N, sm.state + (size_t)group * C::kDim, u};
    float4 *my_trig = sm.trig + group * N;

    const int t0 = blockIdx.y * chunk_len;
    const int t1 = min(tv.n_trans, t0 + chunk_len);
    double acc = 0.0;   // fp64: the sum then depends on how transitions are chunked only at the 1e-16 level
    bool bad_idx = false;
    for (int tb = t0; tb < t1; tb += C::kGroupsPerCta) {       // trip count is CTA-uniform
        const int t = tb + group;
        const bool valid = t < t1;
        const int row = tv.idx ? checked_row(tv.idx[valid ? t : t0], tv.n_rows, bad_idx) : (valid ? t : t0);
        // stage the transition's half-angle pairs, one record per wire
        for (int w = u; w < N; w += C::kGroup) {
            const int f = embed_feature(w, sp.n_feat, sp.feature_map);
            my_trig[w] = f >= 0 ? __ldg(tv.trig + (size_t)row * sp.n_feat + f) : make_float4(1.f, 1.f, 0.f, 0.f);
        }
        group_sync<N>();
        u64 z[kSmemMaxOut];
        run_circuit<N>(cx, z, red_scratch);
        if (u == 0 && valid) {
            const float4 meta = tv.meta[row];
            const int act = __float_as_int(meta.z);
            float qa = 0.f, mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < kSmemMaxOut; ++i) {
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
        group_sync<N>();   // trig/state are reused by the next transition
    }
    const double tot = block_sum_d(bad_idx ? (double)NAN : acc, warp_part);
    if (threadIdx.x == 0) {
        if (gridDim.y == 1) fitness[cand] = (float)(-tot / (double)tv.n_trans);
        else partial[(size_t)cand * gridDim.y + blockIdx.y] = tot;
    }
}

template <int N>
__global__ void __launch_bounds__(Cfg<N>::kThreads)
smem_qvalues_kernel(oqrl_spec sp, const float *__restrict__ params, const float *__restrict__ x, int n_rows,
                    float *__restrict__ q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red_scratch[kSmemMaxOut * 32 * 2];
    using C = Cfg<N>;
    SmemLayout<N> sm(smem_raw, sp.n_layers * N);
    const int cand = blockIdx.x;
    cta_prologue<N>(sp, params + (size_t)cand * sp.n_layers * N * 3, sm);
    __syncthreads();
    const int group = threadIdx.x / C::kGroup, u = threadIdx.x % C::kGroup;
    PassCtx<N> cx{&sp, sm.gates, sm.perm_col, sm.trig + group * N, sm.state + (size_t)group * C::kDim, u};
    float4 *my_trig = sm.trig + group * N;
    const int n_pairs = (n_rows + 1) >> 1;
    for (int jb = blockIdx.y * C::kGroupsPerCta; jb < n_pairs; jb += gridDim.y * C::kGroupsPerCta) {
        const int j = jb + group;
        const bool valid = j < n_pairs;
        const int b0 = valid ? 2 * j : 0, b1 = valid ? min(2 * j + 1, n_rows - 1) : 0;
        for (int w = u; w < N; w += C::kGroup) {
            const int f = embed_feature(w, sp.n_feat, sp.feature_map);
            float4 t = make_float4(1.f, 1.f, 0.f, 0.f);
            if (f >= 0) {
                sincosf(0.5f * x[(size_t)b0 * sp.n_feat + f], &t.z, &t.x);
                sincosf(0.5f * x[(size_t)b1 * sp.n_feat + f], &t.w, &t.y);
            }
            my_trig[w] = t;
        }
        group_sync<N>();
        u64 z[kSmemMaxOut];
        run_circuit<N>(cx, z, red_scratch);
        if (u == 0 && valid) {
#pragma unroll
            for (int i = 0; i < kSmemMaxOut; ++i) {
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
    using C = Cfg<N>;
    const size_t smem = SmemLayout<N>::bytes(sp.n_layers * N);
    if (smem > 227 * 1024) { set_error("n_qubits=%d n_layers=%d needs %zu B of shared memory", N, sp.n_layers, smem); return OQRL_ERR_UNSUPPORTED; }
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(smem_fitness_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int T = tv.n_trans;
    int chunks = 1;
    if (d_partial) {
        const int want = (4 * sm_count + n_cand - 1) / n_cand;
        const int max_chunks = (T + C::kGroupsPerCta - 1) / C::kGroupsPerCta;
        chunks = want < 1 ? 1 : (want > max_chunks ? max_chunks : want);
        if (chunks > 4 * sm_count) chunks = 4 * sm_count;
    }
    int chunk_len = (T + chunks - 1) / chunks;
    chunk_len = ((chunk_len + C::kGroupsPerCta - 1) / C::kGroupsPerCta) * C::kGroupsPerCta;
    chunks = (T + chunk_len - 1) / chunk_len;
    dim3 grid(n_cand, chunks);
    smem_fitness_kernel<N><<<grid, C::kThreads, smem, st>>>(sp, d_params, tv, gamma, chunk_len, d_fitness, d_partial);
    OQRL_LAUNCH_CHECK("smem_fitness_kernel");
    if (chunks > 1) return launch_td_finalize(d_partial, n_cand, chunks, T, d_fitness, st);
    return OQRL_OK;
}

template <int N>
int launch_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q,
                   cudaStream_t st)
{
    using C = Cfg<N>;
    const size_t smem = SmemLayout<N>::bytes(sp.n_layers * N);
    if (smem > 227 * 1024) { set_error("n_qubits=%d n_layers=%d needs %zu B of shared memory", N, sp.n_layers, smem); return OQRL_ERR_UNSUPPORTED; }
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(smem_qvalues_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_pairs = (n_rows + 1) / 2;
    int gy = (n_pairs + C::kGroupsPerCta - 1) / C::kGroupsPerCta;
    if (gy > 2048) gy = 2048;
    dim3 grid(n_cand, gy);
    smem_qvalues_kernel<N><<<grid, C::kThreads, smem, st>>>(sp, d_params, d_x, n_rows, d_q);
    OQRL_LAUNCH_CHECK("smem_qvalues_kernel");
    return OQRL_OK;
}

int check_family(const oqrl_spec &sp)
{
    if (sp.n_layers * sp.n_qubits > kSmemMaxGates) { set_error("shared-memory family supports n_layers*n_qubits <= %d", kSmemMaxGates); return OQRL_ERR_UNSUPPORTED; }
    if (sp.n_out > kSmemMaxOut) { set_error("shared-memory family supports n_out <= %d", kSmemMaxOut); return OQRL_ERR_UNSUPPORTED; }
    return OQRL_OK;
}

#define OQRL_SMEM_DISPATCH(FN, ...)                 \
    switch (sp.n_qubits) {                          \
    case 5: return FN<5>(__VA_ARGS__);              \
    case 6: return FN<6>(__VA_ARGS__);              \
    case 7: return FN<7>(__VA_ARGS__);              \
    case 8: return FN<8>(__VA_ARGS__);              \
    case 9: return FN<9>(__VA_ARGS__);              \
    case 10: return FN<10>(__VA_ARGS__);            \
    case 11: return FN<11>(__VA_ARGS__);            \
    case 12: return FN<12>(__VA_ARGS__);            \
    case 13: return FN<13>(__VA_ARGS__);            \
    default: break;                                 \
    }

}  // namespace

int smem_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
                 float *d_fitness, double *d_partial, int sm_count, cudaStream_t st)
{
    int rc = check_family(sp);
    if (rc) return rc;
    OQRL_SMEM_DISPATCH(launch_fitness, sp, d_params, n_cand, tv, gamma, d_fitness, d_partial, sm_count, st);
    set_error("smem_fitness: n_qubits=%d outside 5..13", sp.n_qubits);
    return OQRL_ERR_INVALID;
}

int smem_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q,
                 cudaStream_t st)
{
    int rc = check_family(sp);
    if (rc) return rc;
    OQRL_SMEM_DISPATCH(launch_qvalues, sp, d_params, n_cand, d_x, n_rows, d_q, st);
    set_error("smem_qvalues: n_qubits=%d outside 5..13", sp.n_qubits);
    return OQRL_ERR_INVALID;
}

// Executed fp32 flop per circuit evaluation.
double smem_flops_per_eval(const oqrl_spec &sp)
{
    const int n = sp.n_qubits, L = sp.n_layers;
    const double dim = (double)(1 << n);
    int n_emb = 0;
    for (int w = 0; w < n; ++w) n_emb += embed_feature(w, sp.n_feat, sp.feature_map) >= 0;
    double f = dim * (1.0 + (n - 4) / 16.0) + dim;    // product state: outer factor per block + 4-level doubling
    f += 14.0 * dim * n * L;                          // SU(2) gates (re-uploaded RY is fused into them)
    if (sp.reupload && L > 1) f += 12.0 * n_emb * (L - 1) * (dim / 16.0);   // per-thread Rot*RY products
    f += 3.0 * dim + (double)sp.n_out * dim;          // |amp|^2 and signed sums
    return f;
}

}  // namespace oqrl
