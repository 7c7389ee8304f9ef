// Specialised 8-qubit PQC kernels (BASELINE config 3: 8 qubits, data re-uploading, large populations).
//
// Same reference code replaced as pqc_smem.cu (/root/reference/src/agent.py:48-61, src/optim.py:64-173); this file
// is the fast path for n = 8, pqc_smem.cu stays the general 5..13-qubit implementation.
//
// Decomposition: a HALF-WARP (16 lanes) owns one transition = the two circuits Q(s), Q(s') packed into fp32x2.
// The 256 amplitudes are split 4 + 4: every lane keeps 16 amplitudes (four "register" qubits, 64 registers) and the
// lane index supplies the other four qubits.  Two layouts alternate inside a layer:
//     layout A: basis index k = (lane << 4) | c   -> wires 4..7 are register qubits
//     layout B: basis index k = (c << 4) | lane   -> wires 0..3 are register qubits
// so a layer is  [4 gates in A] -> transpose through shared memory -> [4 gates in B] -> scatter through shared
// memory with the CNOT ring folded into the addresses (or the CZ ring folded into signs), i.e. two shared-memory
// round trips per layer and NO arithmetic outside registers.  Everything that is not an FFMA2/FMUL2 costs FMA issue
// slots on this part (profiles/r1_summary.md), so the structure is compile-time: gate positions, layouts and the
// swizzled transpose addresses (eight loop-invariant address registers per side, immediate offsets) are fixed, the
// re-upload-fused gate coefficients are computed once per transition by the group (not per lane per gate) and
// layer 0 is folded into the generated product state.
#include "common.cuh"

namespace oqrl {
namespace {

#ifndef OQRL_W8_MIN_CTAS
#define OQRL_W8_MIN_CTAS 4
#endif
constexpr int kW8Threads = 128;
constexpr int kW8Groups = kW8Threads / 16;
constexpr int kW8MaxOut = 4;
constexpr int kW8MaxLayers = 64;

// dynamic shared memory: raw Rot gates, CNOT-ring column tables, per-group wire trigonometry, per-group fused
// coefficient tables, per-group state
struct W8Smem {
    Su2Scalar *rot;      // [L*8]
    unsigned *permcol;   // [7][8]
    float4 *trig;        // [groups][8]   {c_s, c_s', s_s, s_s'} of the wire's feature
    float4 *fused;       // [groups][max(L,1)*8][2]
    float4 *state;       // [groups][256]
    int fused_stride;
    __device__ W8Smem(unsigned char *base, int n_layers)
    {
        const int Lc = n_layers > 0 ? n_layers : 1;
        size_t off = 0;      // state first: the swizzled addressing XORs address bits 4..7 and needs 256-byte alignment
        state = reinterpret_cast<float4 *>(base); off += (size_t)kW8Groups * 256 * sizeof(float4);
        fused = reinterpret_cast<float4 *>(base + off); fused_stride = Lc * 16; off += (size_t)kW8Groups * fused_stride * sizeof(float4);
        trig = reinterpret_cast<float4 *>(base + off); off += (size_t)kW8Groups * 8 * sizeof(float4);
        rot = reinterpret_cast<Su2Scalar *>(base + off); off += (size_t)Lc * 8 * sizeof(Su2Scalar);
        permcol = reinterpret_cast<unsigned *>(base + off);
    }
    static size_t bytes(int n_layers)
    {
        const int Lc = n_layers > 0 ? n_layers : 1;
        return (size_t)Lc * 8 * sizeof(Su2Scalar) + 64 * sizeof(unsigned) + (size_t)kW8Groups * 8 * sizeof(float4)
               + (size_t)kW8Groups * Lc * 16 * sizeof(float4) + (size_t)kW8Groups * 256 * sizeof(float4);
    }
};

struct W8Ctx {
    const oqrl_spec *sp;
    const Su2Scalar *rot;
    const unsigned *permcol;
    const float4 *trig;     // this group's
    float4 *fused;          // this group's
    float4 *state;          // this group's
    int u;                  // lane inside the group (0..15)
    unsigned last_mask;
};

// Group-cooperative: fused coefficient (layers >= 1) or product-state vector (layer 0) of every gate of this transition.
__device__ __forceinline__ void w8_build_fused(const W8Ctx &cx)
{
    const oqrl_spec &sp = *cx.sp;
    const int L = sp.n_layers;
    const int total = (L > 0 ? L : 1) * 8;
    for (int g = cx.u; g < total; g += 16) {
        const int l = g >> 3, w = g & 7;
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
__device__ __forceinline__ void w8_gate(const float4 *coef, u64 (&re)[16], u64 (&im)[16])
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

// Shared-memory slot of basis index k: XOR swizzle so that both the row-wise (layout A) and the column-wise
// (layout B) 128-bit accesses of a quarter-warp hit eight distinct 16-byte bank groups.
__device__ __forceinline__ int w8_slot(int k) { return k ^ ((k >> 4) & 7); }

// Whole circuit for the group's transition.  z[i] (packed s / s') valid in every lane of the group afterwards.
__device__ __forceinline__ void w8_circuit(const W8Ctx &cx, u64 (&z)[kW8MaxOut])
{
    const oqrl_spec &sp = *cx.sp;
    const int L = sp.n_layers, u = cx.u;
    const bool sel = sp.entangler == OQRL_ENT_SEL_CNOT;
    u64 re[16], im[16];

    // ---- generated state in layout A (lane = wires 0..3 = index bits 7..4; registers = wires 4..7 = bits 3..0) ----
    {
        u64 pr = pack2(1.f, 1.f), pi = 0ull;
#pragma unroll
        for (int b = 4; b < 8; ++b) {
            const int wire = 7 - b;
            const float4 v = cx.fused[2 * wire + ((u >> (b - 4)) & 1)];
            const u64 vr = pack2(v.x, v.y), vi = pack2(v.z, v.w);
            const u64 nr = ffma2(fneg2(pi), vi, fmul2(pr, vr));
            pi = ffma2(pi, vr, fmul2(pr, vi));
            pr = nr;
        }
        re[0] = pr; im[0] = pi;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int wire = 7 - q;
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

    // loop-invariant addresses of the two access patterns (bytes into the group's state):
    //   row access   (layout A, k = u<<4 | c):  slot = u<<4 | (c ^ (u & 7))   -> 8 values over c & 7, + 128 B for c & 8
    //   column access (layout B, k = c<<4 | u): slot = c<<4 | (u ^ (c & 7))   -> 8 values over c & 7, + c<<8 immediate
    // The group's state is 4 KB-aligned, so both patterns are "one per-lane base XOR an immediate":
    //   row access    (layout A, k = u<<4 | c): byte = (u<<8 | (u&7)<<4) ^ (c&7)<<4, + 128 for c & 8
    //   column access (layout B, k = c<<4 | u): byte = (u<<4) ^ (c&7)<<4, + c<<8
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(cx.state);
    const unsigned row_base = sbase + ((u << 8) | ((u & 7) << 4));
    const unsigned col_base = sbase + (u << 4);
    auto st128 = [](unsigned addr, u64 a, u64 b) {
        asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
    };
    auto ld128 = [](unsigned addr, u64 &a, u64 &b) {
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
    };
    auto store_rows = [&]() {
#pragma unroll
        for (int c = 0; c < 16; ++c) st128((row_base ^ ((c & 7) << 4)) + ((c & 8) << 4), re[c], im[c]);
    };
    auto load_rows = [&]() {
#pragma unroll
        for (int c = 0; c < 16; ++c) ld128((row_base ^ ((c & 7) << 4)) + ((c & 8) << 4), re[c], im[c]);
    };
    auto load_cols = [&]() {
#pragma unroll
        for (int c = 0; c < 16; ++c) ld128((col_base ^ ((c & 7) << 4)) + (c << 8), re[c], im[c]);
    };
    // Images of the layout's index bits under the layer's CNOT ring M (identity for CZ / no layer):
    // register bit q <-> index bit (cols ? 4 + q : q), lane bit q <-> index bit (cols ? q : 4 + q).
    auto relabel_columns = [&](int layer, bool cols_layout, bool relabel, int (&mc)[4], int &mu) {
        const unsigned *pc = cx.permcol + (relabel ? (sel_range(8, layer) - 1) * 8 : 0);
        mu = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int creg_bit = cols_layout ? 4 + q : q, lane_bit = cols_layout ? q : 4 + q;
            mc[q] = relabel ? (int)pc[creg_bit] : (1 << creg_bit);
            if ((u >> q) & 1) mu ^= relabel ? (int)pc[lane_bit] : (1 << lane_bit);
        }
    };
    // Layer's entangler folded into the write-back.  CNOT ring: element k goes to the slot of M k; the slot function
    // is GF(2)-linear, so the 16 addresses are a Gray-code walk of XORs.  CZ ring: same slot, sign (-1)^{#adjacent 11}.
    auto scatter_entangled = [&](int layer, bool cols_layout) {
        if (sel) {
            int mc[4], mu;
            relabel_columns(layer, cols_layout, true, mc, mu);
            unsigned sm4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) sm4[q] = (unsigned)w8_slot(mc[q]) << 4;
            unsigned o = (unsigned)w8_slot(mu) << 4;
#pragma unroll
            for (int j = 0; j < 16; ++j) {                       // Gray-code order: one XOR per element
                const int c = j ^ (j >> 1);
                if (j > 0) o ^= sm4[(j & 1) ? 0 : ((j & 2) ? 1 : ((j & 4) ? 2 : 3))];
                st128(sbase + o, re[c], im[c]);
            }
        } else {
            // adjacent pairs inside the lane nibble, inside the register nibble (compile time) and the two that cross
            const unsigned par_u = ((u & (u >> 1)) ^ ((u & (u >> 1)) >> 1) ^ ((u & (u >> 1)) >> 2)) & 1u;
            const unsigned u0 = u & 1u, u3 = (u >> 3) & 1u;
            unsigned m[4];   // sign mask by (c0, c3)
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = (par_u ^ ((j & 1) ? u3 : 0u) ^ ((j & 2) ? u0 : 0u)) << 31;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const unsigned cc = c & (c >> 1), cpar = (cc ^ (cc >> 1) ^ (cc >> 2)) & 1u;
                const unsigned mk = m[(c & 1) | ((c >> 2) & 2)] ^ (cpar << 31);
                const u64 mk2 = ((u64)mk << 32) | mk;
                const u64 r = re[c] ^ mk2, i = im[c] ^ mk2;
                if (cols_layout) st128((col_base ^ ((c & 7) << 4)) + (c << 8), r, i);
                else st128((row_base ^ ((c & 7) << 4)) + ((c & 8) << 4), r, i);
            }
        }
    };
    // <Z_i> after the last entangler: signed sum of |amp|^2 with sign = bit (7 - i) of M k.  The sign is linear in
    // the register index, so the 16-term sum is a 4-level butterfly of FMAs with +-1 multipliers.
    auto readout = [&](int layer, bool cols_layout) {
        int mc[4], mu;
        relabel_columns(layer, cols_layout, sel && layer >= 0, mc, mu);
        u64 p[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) p[c] = ffma2(im[c], im[c], fmul2(re[c], re[c]));
#pragma unroll
        for (int i = 0; i < kW8MaxOut; ++i) {
            z[i] = 0ull;
            if (i < sp.n_out) {
                u64 t[8];
                float sg[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) sg[q] = ((mc[q] >> (7 - i)) & 1) ? -1.f : 1.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) t[c] = ffma2(p[2 * c + 1], bcast2(sg[0]), p[2 * c]);
#pragma unroll
                for (int c = 0; c < 4; ++c) t[c] = ffma2(t[2 * c + 1], bcast2(sg[1]), t[2 * c]);
#pragma unroll
                for (int c = 0; c < 2; ++c) t[c] = ffma2(t[2 * c + 1], bcast2(sg[2]), t[2 * c]);
                t[0] = ffma2(t[1], bcast2(sg[3]), t[0]);
                float lo, hi;
                unpack2(t[0], lo, hi);
                if ((mu >> (7 - i)) & 1) { lo = -lo; hi = -hi; }
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    lo += __shfl_xor_sync(0xffffffffu, lo, off);
                    hi += __shfl_xor_sync(0xffffffffu, hi, off);
                }
                z[i] = pack2(lo, hi);
            }
        }
    };

    if (L <= 1) {                       // no further layer: read the generated state out (through layer 0's entangler)
        readout(L == 1 ? 0 : -1, false);
        return;
    }
    scatter_entangled(0, false);
    __syncwarp();
    // One loop iteration = half a layer, so that the unrolled gate code (4 register bits x 128 packed FMAs) exists
    // once: the loop body has to stay inside the instruction caches (the first version, unrolled per layer with
    // per-gate pruning branches, stalled on instruction fetch for 1.9 of every issue cycle: profiles/r1_summary.md).
    const int n_half = 2 * L;
#pragma unroll 1
    for (int h = 2; h < n_half; ++h) {
        const int l = h >> 1;
        const bool second = h & 1;
        if (!second) load_rows();                          // layout A: register bit q <-> wire 7 - q
        else load_cols();                                  // layout B: register bit q <-> index bit 4 + q <-> wire 3 - q
        const float4 *coef = cx.fused + 2 * (l * 8 + (second ? 3 : 7));
        w8_gate<0>(coef, re, im);
        w8_gate<1>(coef - 2, re, im);
        w8_gate<2>(coef - 4, re, im);
        w8_gate<3>(coef - 6, re, im);
        if (!second) {
            store_rows();                                  // same lane, same slots as load_rows: no hazard before it
            __syncwarp();
        } else if (h != n_half - 1) {
            __syncwarp();
            scatter_entangled(l, true);
            __syncwarp();
        }
    }
    readout(L - 1, true);
}

__device__ __forceinline__ void w8_prologue(const oqrl_spec &sp, const float *__restrict__ params, W8Smem &sm)
{
    const int n_gates = sp.n_layers * 8;
    for (int g = threadIdx.x; g < n_gates; g += blockDim.x) {
        Su2Scalar c;
        rot_su2(params[3 * g], params[3 * g + 1], params[3 * g + 2], c.ar, c.ai, c.br, c.bi);
        sm.rot[g] = c;
    }
    for (int e = threadIdx.x; e < 56; e += blockDim.x) sm.permcol[e] = sel_perm(1u << (e & 7), 8, (e >> 3) + 1);
}

__global__ void __launch_bounds__(kW8Threads, OQRL_W8_MIN_CTAS)
w8_fitness_kernel(oqrl_spec sp, const float *__restrict__ params, TransitionView tv, float gamma, int chunk_len,
                  float *__restrict__ fitness, double *__restrict__ partial)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double warp_part[kW8Threads / 32];
    W8Smem sm(smem_raw, sp.n_layers);
    const int cand = blockIdx.x;
    w8_prologue(sp, params + (size_t)cand * sp.n_layers * 24, sm);
    __syncthreads();
    const int group = threadIdx.x >> 4, u = threadIdx.x & 15;
    const unsigned last_mask = (sp.reserved & 1) ? 0xffu : readout_wire_mask(8, sp.n_layers, sp.n_out, sp.entangler);
    W8Ctx cx{&sp, sm.rot, sm.permcol, sm.trig + group * 8, sm.fused + (size_t)group * sm.fused_stride,
             sm.state + (size_t)group * 256, u, last_mask};
    float4 *my_trig = sm.trig + group * 8;

    const int t0 = blockIdx.y * chunk_len;
    const int t1 = min(tv.n_trans, t0 + chunk_len);
    float acc = 0.f;
    for (int tb = t0; tb < t1; tb += kW8Groups) {            // trip count is CTA-uniform
        const int t = tb + group;
        const bool valid = t < t1;
        const int row = tv.idx ? tv.idx[valid ? t : t0] : (valid ? t : t0);
        if (u < 8) {
            const int f = embed_feature(u, sp.n_feat, sp.feature_map);
            my_trig[u] = f >= 0 ? __ldg(tv.trig + (size_t)row * sp.n_feat + f) : make_float4(1.f, 1.f, 0.f, 0.f);
        }
        __syncwarp();
        w8_build_fused(cx);
        __syncwarp();
        u64 z[kW8MaxOut];
        w8_circuit(cx, z);
        if (u == 0 && valid) {
            const float4 meta = tv.meta[row];
            const int act = __float_as_int(meta.z);
            float qa = 0.f, mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < kW8MaxOut; ++i) {
                if (i < sp.n_out) {
                    float lo, hi;
                    unpack2(z[i], lo, hi);
                    if (i == act) qa = lo;
                    mx = fmaxf(mx, hi);
                }
            }
            const float target = meta.x + (1.f - meta.y) * gamma * mx;   // optim.py:163
            const float d = qa - target;
            acc = fmaf(d, d, acc);
        }
        __syncwarp();   // trig / fused / state are reused by the next transition
    }
    double v = (double)acc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kW8Threads / 32; ++w) tot += warp_part[w];
        if (gridDim.y == 1) fitness[cand] = (float)(-tot / (double)tv.n_trans);
        else partial[(size_t)cand * gridDim.y + blockIdx.y] = tot;
    }
}

__global__ void __launch_bounds__(kW8Threads, OQRL_W8_MIN_CTAS)
w8_qvalues_kernel(oqrl_spec sp, const float *__restrict__ params, const float *__restrict__ x, int n_rows, float *__restrict__ q)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    W8Smem sm(smem_raw, sp.n_layers);
    const int cand = blockIdx.x;
    w8_prologue(sp, params + (size_t)cand * sp.n_layers * 24, sm);
    __syncthreads();
    const int group = threadIdx.x >> 4, u = threadIdx.x & 15;
    const unsigned last_mask = (sp.reserved & 1) ? 0xffu : readout_wire_mask(8, sp.n_layers, sp.n_out, sp.entangler);
    W8Ctx cx{&sp, sm.rot, sm.permcol, sm.trig + group * 8, sm.fused + (size_t)group * sm.fused_stride,
             sm.state + (size_t)group * 256, u, last_mask};
    float4 *my_trig = sm.trig + group * 8;
    const int n_pairs = (n_rows + 1) >> 1;
    for (int jb = blockIdx.y * kW8Groups; jb < n_pairs; jb += gridDim.y * kW8Groups) {
        const int j = jb + group;
        const bool valid = j < n_pairs;
        const int b0 = valid ? 2 * j : 0, b1 = valid ? min(2 * j + 1, n_rows - 1) : 0;
        if (u < 8) {
            const int f = embed_feature(u, sp.n_feat, sp.feature_map);
            float4 t = make_float4(1.f, 1.f, 0.f, 0.f);
            if (f >= 0) {
                sincosf(0.5f * x[(size_t)b0 * sp.n_feat + f], &t.z, &t.x);
                sincosf(0.5f * x[(size_t)b1 * sp.n_feat + f], &t.w, &t.y);
            }
            my_trig[u] = t;
        }
        __syncwarp();
        w8_build_fused(cx);
        __syncwarp();
        u64 z[kW8MaxOut];
        w8_circuit(cx, z);
        if (u == 0 && valid) {
#pragma unroll
            for (int i = 0; i < kW8MaxOut; ++i) {
                if (i < sp.n_out) {
                    float lo, hi;
                    unpack2(z[i], lo, hi);
                    q[((size_t)cand * n_rows + b0) * sp.n_out + i] = lo;
                    if (b1 != b0) q[((size_t)cand * n_rows + b1) * sp.n_out + i] = hi;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace

bool w8_supports(const oqrl_spec &sp)
{
    return sp.n_qubits == 8 && sp.n_out <= kW8MaxOut && sp.n_layers <= kW8MaxLayers && W8Smem::bytes(sp.n_layers) <= 56 * 1024;
}

int w8_fitness(const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv, float gamma,
               float *d_fitness, double *d_partial, int sm_count, cudaStream_t st)
{
    const size_t smem = W8Smem::bytes(sp.n_layers);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(w8_fitness_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int T = tv.n_trans;
    int chunks = 1;
    if (d_partial) {
        const int want = (4 * sm_count + n_cand - 1) / n_cand;
        const int max_chunks = (T + kW8Groups - 1) / kW8Groups;
        chunks = want < 1 ? 1 : (want > max_chunks ? max_chunks : want);
        if (chunks > 4 * sm_count) chunks = 4 * sm_count;
    }
    int chunk_len = (T + chunks - 1) / chunks;
    chunk_len = ((chunk_len + kW8Groups - 1) / kW8Groups) * kW8Groups;
    chunks = (T + chunk_len - 1) / chunk_len;
    dim3 grid(n_cand, chunks);
    w8_fitness_kernel<<<grid, kW8Threads, smem, st>>>(sp, d_params, tv, gamma, chunk_len, d_fitness, d_partial);
    OQRL_LAUNCH_CHECK("w8_fitness_kernel");
    if (chunks > 1) return launch_td_finalize(d_partial, n_cand, chunks, T, d_fitness, st);
    return OQRL_OK;
}

int w8_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q, cudaStream_t st)
{
    const size_t smem = W8Smem::bytes(sp.n_layers);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(w8_qvalues_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_pairs = (n_rows + 1) / 2;
    int gy = (n_pairs + kW8Groups - 1) / kW8Groups;
    if (gy > 2048) gy = 2048;
    dim3 grid(n_cand, gy);
    w8_qvalues_kernel<<<grid, kW8Threads, smem, st>>>(sp, d_params, d_x, n_rows, d_q);
    OQRL_LAUNCH_CHECK("w8_qvalues_kernel");
    return OQRL_OK;
}

// Executed fp32 flop per circuit evaluation.
double w8_flops_per_eval(const oqrl_spec &sp)
{
    const int n = 8, L = sp.n_layers;
    const double dim = 256.0;
    int n_emb = 0;
    for (int w = 0; w < n; ++w) n_emb += embed_feature(w, sp.n_feat, sp.feature_map) >= 0;
    double f = 6.0 * 16.0 * (3.0 + 30.0);              // generated state: per lane 3 + 30 complex multiplies
    f += 18.0 * n;                                     // per-wire product vectors (layer 0 folded in)
    if (L > 1) {
        f += 14.0 * dim * n * (L - 1);                 // SU(2) gates, 28 flop per amplitude pair (re-uploaded RY fused in;
                                                       // last-layer gates outside the light cone run as identities)
        if (sp.reupload) f += 12.0 * n_emb * (L - 1);  // one Rot * RY product per gate per transition
    }
    f += 3.0 * dim + (double)sp.n_out * dim;           // |amp|^2 and signed sums
    return f;
}

}  // namespace oqrl
