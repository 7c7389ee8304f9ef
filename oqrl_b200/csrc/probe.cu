// Measurement and test hooks of oqrl_b200 -- a SEPARATE library (liboqrl_b200_probe.so, include/oqrl_b200_probe.h).
// Nothing here is on the product path: bench.py uses the microbenchmarks as roofline denominators, the CPU tests read
// the HBM-class pass list through oqrl_hbm_plan, and both use oqrl_work_model for executed-flop / sweep counts.  The
// library links against liboqrl_b200.so for the planner and the work models (internal C++ symbols of namespace oqrl).
#include <stdlib.h>

#include "common.cuh"
#include "../../include/oqrl_b200_probe.h"

namespace oqrl {

// FP32 FMA-pipe microbenchmark (roofline denominator).
// VARIANT bit 0: 0 = scalar FFMA, 1 = packed FFMA2.  VARIANT >> 1 = operand pattern:
//   0: a = fma(a, m, c)          one fresh register operand per instruction (m, c stay in the reuse cache)
//   1: a[i] = fma(b[i], m, a[i]) two fresh register operands (the pattern of a gate update)
//   2: a[i] = fma(b[i], c[i], a[i]) three fresh register operands
//   3: a[i] = fma(b[i], U, a[i]) like 1 but the multiplier is warp-uniform (kernel parameter)
//   4: a[i] = a[i] * m           multiply only (FMUL / FMUL2), counted as 1 flop per element
//   5: a[i] = a[i] + k           add only (FADD / FADD2), counted as 1 flop per element
//   6: a[i] = fma(b[i], s, a[i]) like 1 with a per-thread 32-bit scalar broadcast to both halves (SASS Rn.F32)
//   7: out of place: a = fma(b, s, c); c = fma(a, s, b); b = fma(c, s, a)  (destination differs from every source,
//      the pattern of a gate update; 3 instructions per chain per round)
template <int VARIANT>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float uni, float *sink)
{
    constexpr int kChains = 16;
    constexpr int PAT = VARIANT >> 1;
    if ((VARIANT & 1) == 0) {
        float a[kChains], b[kChains], c[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            a[i] = seed + (float)(threadIdx.x + i);
            b[i] = 1e-3f * (float)(i + 1) + seed;
            c[i] = 0.5f + 1e-2f * (float)i + seed;
        }
        const float m = 0.999f + seed, k = 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if (PAT == 0) a[i] = fmaf(a[i], m, k);
                if (PAT == 1) a[i] = fmaf(b[i], m, a[i]);
                if (PAT == 2) a[i] = fmaf(b[i], c[i], a[i]);
                if (PAT == 3) a[i] = fmaf(b[i], uni, a[i]);
                if (PAT == 4) a[i] = a[i] * m;
                if (PAT == 5) a[i] = a[i] + k;
                if (PAT == 6) a[i] = fmaf(b[i], c[i & 3], a[i]);
                if (PAT == 7) {
                    const float sb = m + (float)(i & 3) * 1e-7f;
                    a[i] = fmaf(b[i], sb, c[i]);
                    c[i] = fmaf(a[i], sb, b[i]);
                    b[i] = fmaf(c[i], sb, a[i]);
                }
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s += a[i];
        if (s == 123.456f) sink[0] = s;
    } else {
        u64 a[kChains], b[kChains], c[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            a[i] = pack2(seed + (float)(threadIdx.x + i), seed - (float)i);
            b[i] = pack2(1e-3f * (float)(i + 1) + seed, 2e-3f * (float)(i + 1) + seed);
            c[i] = pack2(0.5f + 1e-2f * (float)i + seed, 0.25f + 1e-2f * (float)i + seed);
        }
        const u64 m = pack2(0.999f + seed, 0.998f + seed), k = pack2(1e-3f, 2e-3f), u = pack2(uni, uni);
        float sc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) sc[i] = 1e-6f * (float)(threadIdx.x + i) + seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if (PAT == 0) a[i] = ffma2(a[i], m, k);
                if (PAT == 1) a[i] = ffma2(b[i], m, a[i]);
                if (PAT == 2) a[i] = ffma2(b[i], c[i], a[i]);
                if (PAT == 3) a[i] = ffma2(b[i], u, a[i]);
                if (PAT == 4) a[i] = fmul2(a[i], m);
                if (PAT == 5) a[i] = fadd2(a[i], k);
                if (PAT == 6) a[i] = ffma2(b[i], pack2(sc[i & 3], sc[i & 3]), a[i]);
                if (PAT == 7) {
                    const u64 sb = pack2(sc[i & 3], sc[i & 3]);
                    a[i] = ffma2(b[i], sb, c[i]);
                    c[i] = ffma2(a[i], sb, b[i]);
                    b[i] = ffma2(c[i], sb, a[i]);
                }
                if (PAT == 8) {   // out-of-place multiply: destination differs from both sources
                    const u64 sb = pack2(sc[i & 3], sc[i & 3]);
                    a[i] = fmul2(b[i], sb);
                    b[i] = fmul2(a[i], sb);
                }
                if (PAT == 9) {   // destination = the multiplicand, addend is another register
                    const u64 sb = pack2(sc[i & 3], sc[i & 3]);
                    a[i] = ffma2(sb, a[i], c[i]);
                    c[i] = ffma2(sb, c[i], a[i]);
                }
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kChains; ++i) { float lo, hi; unpack2(a[i], lo, hi); s += lo + hi; }
        if (s == 123.456f) sink[0] = s;
    }
}

// Instruction-fetch probe: a loop whose unrolled body is BODY packed FMAs (16 bytes each) run by warps that start
// at staggered times, i.e. sit at different program counters like the warps of a real kernel.  Used to find how
// large an unrolled hot loop may be before the FMA pipe starves on instruction fetch (scripts/fma_microbench.py).
template <int BODY>
__global__ void __launch_bounds__(128) icache_probe_kernel(int iters, float seed, float uni, float *sink)
{
    u64 a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = pack2(seed + j, seed - j);
    const u64 b = pack2(uni, uni), c = pack2(uni * 0.5f, uni * 0.25f);
    const unsigned wid = (threadIdx.x >> 5) + blockIdx.x * 5u;
    __nanosleep((wid % 7u) * 300u);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < BODY; ++j) a[j & 15] = ffma2(b, c, a[j & 15]);
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { float lo, hi; unpack2(a[j], lo, hi); acc += lo + hi; }
    if (acc == 12345.678f) *sink = acc;
}

template <int BODY>
static void launch_icache_probe(int blocks, int iters, float *sink, cudaStream_t st)
{
    icache_probe_kernel<BODY><<<blocks, 128, 0, st>>>(iters, 0.f, 1e-6f, sink);
}

template <int V>
static void launch_fma_peak(int blocks, int threads, int iters, float *sink, cudaStream_t st)
{
    fma_peak_kernel<V><<<blocks, threads, 0, st>>>(iters, 0.f, 1e-6f, sink);
}


// Gate-body rate: the arithmetic of one register block of the HBM-resident class (16 amplitudes packed (re, im),
// four SU(2) gates = 256 FFMA2-class instructions) repeated on registers, no memory traffic.  form 0 = the
// complex-packed body of pqc_hbm.cu (half-swap / per-half negation operand modifiers); form 1 = the same gates as
// scalar FFMA; form 2 = two-circuit packing (the register / shared-memory families: 16 packed ops per two pairs).
template <int FORM>
__global__ void __launch_bounds__(128) gate_rate_kernel(int iters, const float4 *__restrict__ coef, float2 *__restrict__ sink)
{
    if (FORM == 0) {
        GateOps ops[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ops[i] = load_gate_ops(coef + 2 * i);
        u64 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = pack2(1e-3f * (float)(threadIdx.x + m), 2e-3f * (float)m);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    if (((m >> i) & 1) == 0) su2_cpair(ops[i], v[m], v[m | (1 << i)]);
        }
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int m = 0; m < 16; ++m) { float a, b; unpack2(v[m], a, b); re += a; im += b; }
        if (re == 123.456f) sink[0] = make_float2(re, im);
    } else if (FORM == 1) {
        float ar[4], ai[4], br[4], bi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float4 a = coef[2 * i], b = coef[2 * i + 1]; ar[i] = a.x; br[i] = a.y; ai[i] = b.x; bi[i] = b.z; }
        float xr[16], xi[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) { xr[m] = 1e-3f * (float)(threadIdx.x + m); xi[m] = 2e-3f * (float)m; }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    if (((m >> i) & 1) == 0) {
                        const int p = m | (1 << i);
                        const float x0r = xr[m], x0i = xi[m], x1r = xr[p], x1i = xi[p];
                        float n0r = ar[i] * x0r, n0i = ar[i] * x0i, n1r = -br[i] * x0r, n1i = -br[i] * x0i;
                        n0r = fmaf(-ai[i], x0i, n0r); n0i = fmaf(ai[i], x0r, n0i); n1r = fmaf(-bi[i], x0i, n1r); n1i = fmaf(bi[i], x0r, n1i);
                        n0r = fmaf(br[i], x1r, n0r); n0i = fmaf(br[i], x1i, n0i); n1r = fmaf(ar[i], x1r, n1r); n1i = fmaf(ar[i], x1i, n1i);
                        n0r = fmaf(-bi[i], x1i, n0r); n0i = fmaf(bi[i], x1r, n0i); n1r = fmaf(ai[i], x1i, n1r); n1i = fmaf(-ai[i], x1r, n1i);
                        xr[m] = n0r; xi[m] = n0i; xr[p] = n1r; xi[p] = n1i;
                    }
        }
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int m = 0; m < 16; ++m) { re += xr[m]; im += xi[m]; }
        if (re == 123.456f) sink[0] = make_float2(re, im);
    } else {
        Su2Scalar g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float4 a = coef[2 * i], b = coef[2 * i + 1]; g[i].ar = a.x; g[i].br = a.y; g[i].ai = b.x; g[i].bi = b.z; }
        u64 re[16], im[16];       // two circuits per value: 16 amplitudes x 2 circuits
#pragma unroll
        for (int m = 0; m < 16; ++m) { re[m] = pack2(1e-3f * (float)(threadIdx.x + m), 3e-3f * (float)m); im[m] = pack2(2e-3f * (float)m, 1e-3f); }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    if (((m >> i) & 1) == 0) su2_pair(g[i], re[m], im[m], re[m | (1 << i)], im[m | (1 << i)]);
        }
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int m = 0; m < 16; ++m) { float lo, hi; unpack2(re[m], lo, hi); a += lo + hi; unpack2(im[m], lo, hi); b += lo + hi; }
        if (a == 123.456f) sink[0] = make_float2(a, b);
    }
}

// Register-block body WITH its shared-memory traffic (forms 3..7): each thread loads 16 amplitudes from a 64 KiB
// buffer, applies four gates, stores them back -- what one group of the HBM pass kernel does to a tile, minus TMA and
// barriers.  VARIANT 0: asm volatile ld/st with a memory clobber (the kernel's ld_amp / st_amp); 1: plain C++ shared
// pointer accesses; 2: as 0 with two blocks streamed through two register sets; 3: as 0, addresses XOR-composed per
// element like the kernel's (rel0 ^ off[m]); 4: as 1 with two blocks streamed.
template <int VARIANT, bool VECTOR_COEF = false>
__global__ void __launch_bounds__(128) gate_smem_kernel(int iters, const float4 *__restrict__ coef, float2 *__restrict__ sink, unsigned xor_seed)
{
    extern __shared__ __align__(16) unsigned char smem[];
    GateOps ops[4];
    // VECTOR_COEF: the address is not provably warp-uniform (it is, at run time), so the coefficients stay in vector
    // registers and every FFMA2 reads three 64-bit vector operands -- instead of uniform-register coefficients
#pragma unroll
    for (int i = 0; i < 4; ++i) ops[i] = load_gate_ops(coef + 2 * i + (VECTOR_COEF ? (threadIdx.x & xor_seed) : 0));
    u64 *buf = reinterpret_cast<u64 *>(smem);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = pack2(1e-3f * (float)(i & 255), 2e-3f * (float)(i >> 8));
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    auto ld = [&](unsigned a) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; };
    auto st = [&](unsigned a, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); };
    // block b of a thread: elements at amplitude index  (b * 128 + tid) + 512 * m'  with m' a permutation of m (XOR seed)
    unsigned off[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) off[m] = (unsigned)(m * 512) * 8u;
    if (VARIANT == 3) {
        const unsigned d[4] = {(512u ^ (xor_seed & 0u)) * 8u, 1024u * 8u, 2048u * 8u, 4096u * 8u};
        off[0] = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < (1 << i); ++m) off[m | (1 << i)] = off[m] ^ d[i];
    }
    auto gates = [&](u64 (&v)[16]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < 16; ++m)
                if (((m >> i) & 1) == 0) su2_cpair(ops[i], v[m], v[m | (1 << i)]);
    };
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (VARIANT == 0 || VARIANT == 3) {
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const unsigned rel = (unsigned)(b * 128 + threadIdx.x) * 8u;
                u64 v[16];
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = ld(sbase + (VARIANT == 3 ? (rel ^ off[m]) : (rel + off[m])));
                gates(v);
#pragma unroll
                for (int m = 0; m < 16; ++m) st(sbase + (VARIANT == 3 ? (rel ^ off[m]) : (rel + off[m])), v[m]);
            }
        } else if (VARIANT == 1) {
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                u64 *p = buf + b * 128 + threadIdx.x;
                u64 v[16];
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = p[m * 512];
                gates(v);
#pragma unroll
                for (int m = 0; m < 16; ++m) p[m * 512] = v[m];
            }
        } else if (VARIANT == 2) {
            u64 va[16], vb[16];
            unsigned ra = (unsigned)threadIdx.x * 8u, rb = (unsigned)(128 + threadIdx.x) * 8u;
#pragma unroll
            for (int m = 0; m < 16; ++m) va[m] = ld(sbase + ra + off[m]);
#pragma unroll
            for (int m = 0; m < 16; ++m) vb[m] = ld(sbase + rb + off[m]);
#pragma unroll 1
            for (int b = 0; b < 4; b += 2) {
                gates(va);
#pragma unroll
                for (int m = 0; m < 16; ++m) st(sbase + ra + off[m], va[m]);
                if (b + 2 < 4) {
                    ra = (unsigned)((b + 2) * 128 + threadIdx.x) * 8u;
#pragma unroll
                    for (int m = 0; m < 16; ++m) va[m] = ld(sbase + ra + off[m]);
                }
                gates(vb);
#pragma unroll
                for (int m = 0; m < 16; ++m) st(sbase + rb + off[m], vb[m]);
                if (b + 3 < 4) {
                    rb = (unsigned)((b + 3) * 128 + threadIdx.x) * 8u;
#pragma unroll
                    for (int m = 0; m < 16; ++m) vb[m] = ld(sbase + rb + off[m]);
                }
            }
        } else {
            u64 va[16], vb[16];
            u64 *pa = buf + threadIdx.x, *pb = buf + 128 + threadIdx.x;
#pragma unroll
            for (int m = 0; m < 16; ++m) va[m] = pa[m * 512];
#pragma unroll
            for (int m = 0; m < 16; ++m) vb[m] = pb[m * 512];
#pragma unroll 1
            for (int b = 0; b < 4; b += 2) {
                gates(va);
#pragma unroll
                for (int m = 0; m < 16; ++m) pa[m * 512] = va[m];
                if (b + 2 < 4) {
                    pa = buf + (b + 2) * 128 + threadIdx.x;
#pragma unroll
                    for (int m = 0; m < 16; ++m) va[m] = pa[m * 512];
                }
                gates(vb);
#pragma unroll
                for (int m = 0; m < 16; ++m) pb[m * 512] = vb[m];
                if (b + 3 < 4) {
                    pb = buf + (b + 3) * 128 + threadIdx.x;
#pragma unroll
                    for (int m = 0; m < 16; ++m) vb[m] = pb[m * 512];
                }
            }
        }
    }
    __syncthreads();
    const u64 r = buf[threadIdx.x];
    float a, b;
    unpack2(r, a, b);
    if (a == 123.456f) sink[0] = make_float2(a, b);
}

// Form 9: the same register block with its Rot gates split as RZ(omega) RY(theta) RZ(phi).  The four RZ(phi) of a block
// are ONE diagonal (a 16-entry phase table indexed by the element's role bits), likewise the four RZ(omega); between
// them four real RY gates.  Per amplitude: 2 complex multiplies by table entries (2 packed ops each) + 4 real pair
// updates (2 packed ops per amplitude each) = 12 packed ops instead of 16, and only the table multiplies read a
// 64-bit coefficient pair.  Table entries are 16 bytes {t_re, -, t_im, t_im}, loaded per block with LDS.128 from an
// address the compiler cannot prove uniform (as in the pass kernel, where it depends on the circuit).
__global__ void __launch_bounds__(128) gate_smem_zyz_kernel(int iters, const float4 *__restrict__ coef, float2 *__restrict__ sink, unsigned xor_seed)
{
    extern __shared__ __align__(16) unsigned char smem[];
    u64 *buf = reinterpret_cast<u64 *>(smem);
    float4 *tab = reinterpret_cast<float4 *>(smem + 65536);            // [2][16]
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = pack2(1e-3f * (float)(i & 255), 2e-3f * (float)(i >> 8));
    if (threadIdx.x < 32) {
        const float a = 0.05f * (float)(threadIdx.x + 1);
        tab[threadIdx.x] = make_float4(cosf(a), 0.f, sinf(a), sinf(a));
    }
    float c[4], sn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float4 r = coef[2 * i + (threadIdx.x & xor_seed)]; c[i] = r.x; sn[i] = r.y; }
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned tbase = sbase + 65536u + (threadIdx.x & xor_seed) * 16u;
    auto ld = [&](unsigned a) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; };
    auto st = [&](unsigned a, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); };
    auto ldt = [&](unsigned a, float &tr, u64 &pti) {
        unsigned r0, r1; u64 p;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(p), "=l"(pti) : "r"(a) : "memory");
        float lo, hi; unpack2(p, lo, hi); tr = lo; (void)hi; (void)r0; (void)r1;
    };
    unsigned off[16];
    {
        const unsigned d[4] = {(512u ^ (xor_seed & 0u)) * 8u, 1024u * 8u, 2048u * 8u, 4096u * 8u};
        off[0] = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < (1 << i); ++m) off[m | (1 << i)] = off[m] ^ d[i];
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const unsigned rel = (unsigned)(b * 128 + threadIdx.x) * 8u;
            u64 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = ld(sbase + (rel ^ off[m]));
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                float tr; u64 pti;
                ldt(tbase + m * 16u, tr, pti);
                v[m] = ffma2(neg_lo(pti), cswap(v[m]), fmul2(pack2(tr, tr), v[m]));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    if (((m >> i) & 1) == 0) {
                        const u64 x0 = v[m], x1 = v[m | (1 << i)];
                        const u64 cc = pack2(c[i], c[i]), ss = pack2(sn[i], sn[i]), ns = pack2(-sn[i], -sn[i]);
                        v[m] = ffma2(ns, x1, fmul2(cc, x0));
                        v[m | (1 << i)] = ffma2(ss, x0, fmul2(cc, x1));
                    }
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                float tr; u64 pti;
                ldt(tbase + 256u + m * 16u, tr, pti);
                v[m] = ffma2(neg_lo(pti), cswap(v[m]), fmul2(pack2(tr, tr), v[m]));
            }
#pragma unroll
            for (int m = 0; m < 16; ++m) st(sbase + (rel ^ off[m]), v[m]);
        }
    }
    __syncthreads();
    const u64 r = buf[threadIdx.x];
    float a, b;
    unpack2(r, a, b);
    if (a == 123.456f) sink[0] = make_float2(a, b);
}

// ---- where should the state of a 16-qubit circuit live between its two passes? ------------------------------------
// The exchange between the pass-1 tiling and the pass-2 tiling of one 2^16-amplitude state (512 KiB, eight 64 KiB
// tiles): every tile's 256 rows (256 bytes each) spread evenly over the eight tiles of the next pass.  Measured in the
// two media a B200 offers, with the same access pattern and the same copy engine (cp.async.bulk):
//   MEDIUM 0: distributed shared memory -- a cluster of eight CTAs holds the state; each CTA pushes its rows into its
//             peers' landing buffers (cp.async.bulk.shared::cluster.shared::cta, completion on the peer's mbarrier);
//   MEDIUM 1: L2 / HBM -- each CTA bulk-stores its rows to a global scratch state and, once the cluster has stored,
//             bulk-loads the rows of its next tile (what the production pass kernel does, one launch per pass);
//   MEDIUM 2: as 0 with plain st.shared::cluster stores by all threads instead of bulk copies.
// Nothing overlaps the exchange here: it is the exposed cost of a transposition, per circuit and cluster.
constexpr int kXRows = 256, kXRowBytes = 256, kXTile = kXRows * kXRowBytes;
__device__ __forceinline__ uint32_t x_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void x_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool x_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
template <int MEDIUM>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(128) exchange_kernel(int iters, float2 *scratch, unsigned *fault)
{
    extern __shared__ __align__(1024) unsigned char xs[];
    unsigned char *src = xs, *land = xs + kXTile;
    __shared__ unsigned long long bar;
    unsigned rank, cluster_id;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cluster_id));
    for (int i = threadIdx.x; i < kXTile / 16; i += blockDim.x) reinterpret_cast<float4 *>(src)[i] = make_float4((float)i, (float)rank, 1.f, 2.f);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(x_smem(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    x_cluster_sync();
    unsigned char *gstate = reinterpret_cast<unsigned char *>(scratch) + (size_t)cluster_id * 8 * kXTile;
    for (int it = 0; it < iters; ++it) {
        // row r of this tile belongs to tile (r & 7) of the next pass, where it becomes row (rank * 32 + (r >> 3))
        if (MEDIUM == 0) {
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(x_smem(&bar)), "r"(kXTile) : "memory");
            x_cluster_sync();                                       // every peer's barrier is armed, its landing buffer is free
            for (int r = threadIdx.x; r < kXRows; r += blockDim.x) {
                const unsigned dst_cta = r & 7, dst_row = rank * 32 + (r >> 3);
                uint32_t rdst, rbar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(x_smem(land + dst_row * kXRowBytes)), "r"(dst_cta));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(x_smem(&bar)), "r"(dst_cta));
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(rdst), "r"(x_smem(src + r * kXRowBytes)), "r"(kXRowBytes), "r"(rbar) : "memory");
            }
            const long long t0 = clock64();
            while (!x_try_wait(x_smem(&bar), it & 1))
                if (clock64() - t0 > (2LL << 30)) { *fault = 1u; __trap(); }
        } else if (MEDIUM == 2) {
            x_cluster_sync();
            for (int i = threadIdx.x; i < kXTile / 16; i += blockDim.x) {
                const int r = i >> 4, c = i & 15;
                const unsigned dst_cta = r & 7, dst_row = rank * 32 + (r >> 3);
                uint32_t rdst;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(x_smem(land + dst_row * kXRowBytes + c * 16)), "r"(dst_cta));
                const float4 v = reinterpret_cast<const float4 *>(src)[i];
                asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rdst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
            x_cluster_sync();                                       // all stores have landed
        } else {
            for (int r = threadIdx.x; r < kXRows; r += blockDim.x) {
                const unsigned dst_cta = r & 7, dst_row = rank * 32 + (r >> 3);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(gstate + ((size_t)dst_cta * kXRows + dst_row) * kXRowBytes), "r"(x_smem(src + r * kXRowBytes)), "r"(kXRowBytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(x_smem(&bar)), "r"(kXTile) : "memory");
            x_cluster_sync();                                       // the whole state is in global memory
            for (int r = threadIdx.x; r < kXRows; r += blockDim.x)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(x_smem(land + r * kXRowBytes)), "l"(gstate + ((size_t)rank * kXRows + r) * kXRowBytes), "r"(kXRowBytes), "r"(x_smem(&bar)) : "memory");
            const long long t0 = clock64();
            while (!x_try_wait(x_smem(&bar), it & 1))
                if (clock64() - t0 > (2LL << 30)) { *fault = 1u; __trap(); }
        }
        __syncthreads();
        // the landed tile becomes the source of the next exchange
        unsigned char *t = src; src = land; land = t;
    }
    x_cluster_sync();
    if (reinterpret_cast<float4 *>(src)[threadIdx.x].z == 123.f) scratch[0] = make_float2(1.f, 1.f);
}

}  // namespace oqrl

using namespace oqrl;

extern "C" {

int oqrl_work_model(const oqrl_spec *spec, double *flops_per_eval, double *state_sweeps)
{
    return work_model(spec, flops_per_eval, state_sweeps);
}

int oqrl_reg_plan(int32_t n_cand, int32_t n_trans, int32_t n_warps, int32_t can_split, int64_t partial_capacity, int32_t *out7)
{
    if (!out7 || n_cand < 0 || n_trans <= 0 || n_warps <= 0) { set_error("bad reg_plan arguments"); return OQRL_ERR_INVALID; }
    reg_plan_dump(n_cand, n_trans, n_warps, can_split, (long long)partial_capacity, out7);
    return OQRL_OK;
}

int oqrl_hbm_plan(const oqrl_spec *spec, void *buf, uint64_t capacity, uint64_t *n_bytes, int32_t keep_state)
{
    int rc = work_model(spec, nullptr, nullptr);     // validates the spec
    if (rc) return rc;
    size_t nb = 0;
    rc = hbm_plan_dump(*spec, buf, (size_t)capacity, &nb, keep_state);
    if (n_bytes) *n_bytes = (uint64_t)nb;
    return rc;
}

int oqrl_fma_peak(int32_t variant, int32_t iters, double *tflops, void *stream)
{
    // bits 8..15 of `variant`: resident warps per SM sub-partition to run with (0 = 16, the maximum)
    const int warps_per_smsp = (variant >> 8) & 0xff;
    variant &= 0xff;
    if (!tflops || iters <= 0 || variant < 0 || (variant > 19 && (variant < 32 || variant > 39)) || warps_per_smsp > 16) { set_error("bad fma_peak arguments"); return OQRL_ERR_INVALID; }
    static const int kProbeBody[8] = {256, 384, 512, 768, 1024, 1536, 2048, 3072};   // variants 32..39: instruction-fetch probe
    int sm_count = 0;
    int rc = device_sm_count(&sm_count);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float *sink;
    OQRL_CUDA_CHECK(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    OQRL_CUDA_CHECK(cudaEventCreate(&e0));
    OQRL_CUDA_CHECK(cudaEventCreate(&e1));
    int blocks = sm_count * 8, threads = 256;
    if (warps_per_smsp) { threads = 128; blocks = sm_count * warps_per_smsp; }   // one warp per sub-partition per CTA
    if (variant >= 32) { threads = 128; blocks = sm_count * (warps_per_smsp ? warps_per_smsp : 4); }
    for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
        OQRL_CUDA_CHECK(cudaEventRecord(e0, st));
        switch (variant) {
        case 0: launch_fma_peak<0>(blocks, threads, iters, sink, st); break;
        case 1: launch_fma_peak<1>(blocks, threads, iters, sink, st); break;
        case 2: launch_fma_peak<2>(blocks, threads, iters, sink, st); break;
        case 3: launch_fma_peak<3>(blocks, threads, iters, sink, st); break;
        case 4: launch_fma_peak<4>(blocks, threads, iters, sink, st); break;
        case 5: launch_fma_peak<5>(blocks, threads, iters, sink, st); break;
        case 6: launch_fma_peak<6>(blocks, threads, iters, sink, st); break;
        case 7: launch_fma_peak<7>(blocks, threads, iters, sink, st); break;
        case 8: launch_fma_peak<8>(blocks, threads, iters, sink, st); break;
        case 9: launch_fma_peak<9>(blocks, threads, iters, sink, st); break;
        case 10: launch_fma_peak<10>(blocks, threads, iters, sink, st); break;
        case 11: launch_fma_peak<11>(blocks, threads, iters, sink, st); break;
        case 12: launch_fma_peak<12>(blocks, threads, iters, sink, st); break;
        case 13: launch_fma_peak<13>(blocks, threads, iters, sink, st); break;
        case 14: launch_fma_peak<14>(blocks, threads, iters, sink, st); break;
        case 15: launch_fma_peak<15>(blocks, threads, iters, sink, st); break;
        case 17: launch_fma_peak<17>(blocks, threads, iters, sink, st); break;
        case 19: launch_fma_peak<19>(blocks, threads, iters, sink, st); break;
        case 32: launch_icache_probe<256>(blocks, iters, sink, st); break;
        case 33: launch_icache_probe<384>(blocks, iters, sink, st); break;
        case 34: launch_icache_probe<512>(blocks, iters, sink, st); break;
        case 35: launch_icache_probe<768>(blocks, iters, sink, st); break;
        case 36: launch_icache_probe<1024>(blocks, iters, sink, st); break;
        case 37: launch_icache_probe<1536>(blocks, iters, sink, st); break;
        case 38: launch_icache_probe<2048>(blocks, iters, sink, st); break;
        case 39: launch_icache_probe<3072>(blocks, iters, sink, st); break;
        default: set_error("variant %d has no scalar counterpart", variant); cudaFree(sink); return OQRL_ERR_INVALID;
        }
        OQRL_LAUNCH_CHECK("fma_peak_kernel");
        OQRL_CUDA_CHECK(cudaEventRecord(e1, st));
        OQRL_CUDA_CHECK(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    OQRL_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (variant >= 32) {
        *tflops = (double)blocks * threads * (double)iters * kProbeBody[variant - 32] * 4.0 / ((double)ms * 1e-3) / 1e12;
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
        return OQRL_OK;
    }
    const double flop = (double)blocks * threads * (double)iters * 16.0 * (((variant >> 1) == 4 || (variant >> 1) == 5) ? 1.0 : ((variant >> 1) == 7 ? 6.0 : ((variant >> 1) == 8 ? 2.0 : ((variant >> 1) == 9 ? 4.0 : 2.0)))) * ((variant & 1) ? 2.0 : 1.0);
    *tflops = flop / ((double)ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return OQRL_OK;
}


int oqrl_gate_rate(int32_t form, int32_t warps_per_smsp, int32_t iters, double *tflops, void *stream)
{
    if (!tflops || form < 0 || form > 9 || warps_per_smsp < 1 || warps_per_smsp > 12 || iters <= 0) { set_error("bad gate_rate arguments"); return OQRL_ERR_INVALID; }
    int sm_count = 0;
    int rc = device_sm_count(&sm_count);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float4 h_coef[8];
    for (int i = 0; i < 4; ++i) {      // a = cos t e^{i p}, b = sin t e^{i q}: unitary, values stay bounded
        const float t = 0.3f + 0.2f * (float)i, p = 0.1f * (float)(i + 1), q = 0.7f - 0.15f * (float)i;
        const float ar = cosf(t) * cosf(p), ai = cosf(t) * sinf(p), br = sinf(t) * cosf(q), bi = sinf(t) * sinf(q);
        h_coef[2 * i] = make_float4(ar, br, 0.f, 0.f);
        h_coef[2 * i + 1] = make_float4(ai, ai, bi, bi);
    }
    float4 *d_coef;
    float2 *sink;
    OQRL_CUDA_CHECK(cudaMalloc(&d_coef, sizeof(h_coef)));
    OQRL_CUDA_CHECK(cudaMalloc(&sink, sizeof(float2)));
    OQRL_CUDA_CHECK(cudaMemcpyAsync(d_coef, h_coef, sizeof(h_coef), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    OQRL_CUDA_CHECK(cudaEventCreate(&e0));
    OQRL_CUDA_CHECK(cudaEventCreate(&e1));
    const int blocks = sm_count * warps_per_smsp;       // 128-thread CTAs: one warp per sub-partition each
    for (int rep = 0; rep < 2; ++rep) {
        OQRL_CUDA_CHECK(cudaEventRecord(e0, st));
        if (form == 0) gate_rate_kernel<0><<<blocks, 128, 0, st>>>(iters, d_coef, sink);
        else if (form == 1) gate_rate_kernel<1><<<blocks, 128, 0, st>>>(iters, d_coef, sink);
        else if (form == 2) gate_rate_kernel<2><<<blocks, 128, 0, st>>>(iters, d_coef, sink);
        else {
            // forms 3..7: one CTA of 128 threads per SM (64 KiB of shared memory each), `warps_per_smsp` ignored above 1:
            // CTAs per SM = warps_per_smsp (each brings its own buffer)
            const int smem = 65536;
            const int iters_s = iters / 4 > 0 ? iters / 4 : 1;       // four blocks per iteration
            switch (form) {
            case 3: cudaFuncSetAttribute(gate_smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); gate_smem_kernel<0><<<blocks, 128, smem, st>>>(iters_s, d_coef, sink, 0u); break;
            case 4: cudaFuncSetAttribute(gate_smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); gate_smem_kernel<1><<<blocks, 128, smem, st>>>(iters_s, d_coef, sink, 0u); break;
            case 5: cudaFuncSetAttribute(gate_smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); gate_smem_kernel<2><<<blocks, 128, smem, st>>>(iters_s, d_coef, sink, 0u); break;
            case 6: cudaFuncSetAttribute(gate_smem_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); gate_smem_kernel<3><<<blocks, 128, smem, st>>>(iters_s, d_coef, sink, 0u); break;
            case 7: cudaFuncSetAttribute(gate_smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); gate_smem_kernel<4><<<blocks, 128, smem, st>>>(iters_s, d_coef, sink, 0u); break;
            case 9: cudaFuncSetAttribute(gate_smem_zyz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + 512); gate_smem_zyz_kernel<<<blocks, 128, smem + 512, st>>>(iters_s, d_coef, sink, 0u); break;
            default: cudaFuncSetAttribute(gate_smem_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); gate_smem_kernel<3, true><<<blocks, 128, smem, st>>>(iters_s, d_coef, sink, 0u); break;
            }
        }
        OQRL_LAUNCH_CHECK("gate_rate_kernel");
        OQRL_CUDA_CHECK(cudaEventRecord(e1, st));
        OQRL_CUDA_CHECK(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    OQRL_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    // 28 flop per amplitude pair and gate; forms 0/1: 8 pairs x 4 gates per iteration, form 2: twice that (two circuits)
    const double blocks_done = form >= 3 ? (double)(iters / 4 > 0 ? iters / 4 : 1) * 4.0 : (double)iters;
    const double flop = (double)blocks * 128.0 * blocks_done * 4.0 * 8.0 * 28.0 * (form == 2 ? 2.0 : 1.0);
    *tflops = flop / ((double)ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_coef); cudaFree(sink);
    return OQRL_OK;
}

int oqrl_exchange_rate(int32_t medium, int32_t iters, double *us_per_exchange, double *gbytes_per_s, void *stream)
{
    if (!us_per_exchange || medium < 0 || medium > 2 || iters <= 0) { set_error("bad exchange_rate arguments"); return OQRL_ERR_INVALID; }
    int sm_count = 0;
    int rc = device_sm_count(&sm_count);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int clusters = sm_count / 8;                       // 18 clusters of eight CTAs on 148 SMs
    float2 *scratch;
    unsigned *fault;
    OQRL_CUDA_CHECK(cudaMalloc(&scratch, (size_t)clusters * 8 * kXTile));
    OQRL_CUDA_CHECK(cudaMalloc(&fault, sizeof(unsigned)));
    OQRL_CUDA_CHECK(cudaMemsetAsync(fault, 0, sizeof(unsigned), st));
    const int smem = 2 * kXTile;
    void (*kern)(int, float2 *, unsigned *) = medium == 0 ? exchange_kernel<0> : (medium == 1 ? exchange_kernel<1> : exchange_kernel<2>);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    OQRL_CUDA_CHECK(cudaEventCreate(&e0));
    OQRL_CUDA_CHECK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
        OQRL_CUDA_CHECK(cudaEventRecord(e0, st));
        kern<<<clusters * 8, 128, smem, st>>>(iters, scratch, fault);
        OQRL_LAUNCH_CHECK("exchange_kernel");
        OQRL_CUDA_CHECK(cudaEventRecord(e1, st));
        OQRL_CUDA_CHECK(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    OQRL_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *us_per_exchange = (double)ms * 1e3 / (double)iters;
    if (gbytes_per_s) *gbytes_per_s = (double)clusters * 8.0 * kXTile * (double)iters / ((double)ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(scratch); cudaFree(fault);
    return OQRL_OK;
}

}  // extern "C"
