// C-ABI entry points of liboqrl_b200 (include/oqrl_b200.h): argument checking,
// dataset residency, kernel-family dispatch and the small shared kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>

#include "common.cuh"

namespace oqrl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- device bookkeeping -------------------------------------------------------
struct DeviceCtx {
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    void *workspace = nullptr;   // grown on demand, reused across calls on this device
    size_t workspace_bytes = 0;
    // cross-chunk reduction state of the register family: per-candidate partial sums and
    // arrival counters (zero between launches: the kernel resets the ones it used)
    double *red_partial = nullptr;
    unsigned *red_arrivals = nullptr;
    int red_capacity = 0;
    long long red_doubles = 0;
    // persistent staging area of the host-buffer entry points (a stream-ordered allocation per call would be
    // returned to the OS at every synchronisation with the default pool settings)
    char *staging = nullptr;
    size_t staging_bytes = 0;
};
static std::mutex g_ctx_mu;
static DeviceCtx g_ctx[16];

static int current_ctx(DeviceCtx **out)
{
    int dev = 0;
    OQRL_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) { set_error("device ordinal %d out of range", dev); return OQRL_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    DeviceCtx &c = g_ctx[dev];
    if (c.device < 0) {
        cudaDeviceProp p;
        OQRL_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
        if (p.major != 10) {
            set_error("oqrl_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
            return OQRL_ERR_UNSUPPORTED;
        }
        c.device = dev; c.sm_count = p.multiProcessorCount; c.cc_major = p.major; c.cc_minor = p.minor;
    }
    *out = &c;
    return OQRL_OK;
}

static int ensure_workspace(DeviceCtx *c, size_t bytes, cudaStream_t st, void **out)
{
    if (bytes > c->workspace_bytes) {
        if (c->workspace) {
            OQRL_CUDA_CHECK(cudaStreamSynchronize(st));
            OQRL_CUDA_CHECK(cudaFree(c->workspace));
            c->workspace = nullptr; c->workspace_bytes = 0;
        }
        size_t want = bytes + bytes / 8 + (1 << 20);
        cudaError_t e = cudaMalloc(&c->workspace, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return OQRL_ERR_NOMEM;
        }
        c->workspace_bytes = want;
    }
    *out = c->workspace;
    return OQRL_OK;
}

static int ensure_staging(DeviceCtx *c, size_t bytes, cudaStream_t st, char **out)
{
    if (bytes > c->staging_bytes) {
        if (c->staging) {
            OQRL_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(c->staging);
            c->staging = nullptr; c->staging_bytes = 0;
        }
        const size_t want = bytes + bytes / 4 + (64 << 10);
        cudaError_t e = cudaMalloc((void **)&c->staging, want);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("staging allocation of %zu bytes failed: %s", want, cudaGetErrorString(e)); return OQRL_ERR_NOMEM; }
        c->staging_bytes = want;
    }
    *out = c->staging;
    return OQRL_OK;
}

// scratch of the register family's split work items: one fp64 sum per (candidate, 256-transition segment) and one
// arrival counter per candidate
static int ensure_reduce_buffers(DeviceCtx *c, int n_cand, long long n_segs, cudaStream_t st)
{
    long long want = (long long)n_cand * n_segs;
    if (want > (64LL << 20) / 8) want = (64LL << 20) / 8;            // beyond 64 MiB the launcher simply does not split
    if (want < (long long)n_cand * 4) want = (long long)n_cand * 4;
    if (n_cand <= c->red_capacity && want <= c->red_doubles) return OQRL_OK;
    if (c->red_partial) {
        OQRL_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(c->red_partial); cudaFree(c->red_arrivals);
        c->red_partial = nullptr; c->red_arrivals = nullptr; c->red_capacity = 0; c->red_doubles = 0;
    }
    const int cap = n_cand + n_cand / 4 + 256;
    const long long doubles = want + want / 4 + 1024;
    cudaError_t e = cudaMalloc((void **)&c->red_partial, (size_t)doubles * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->red_arrivals, (size_t)cap * sizeof(unsigned));
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(c->red_partial); c->red_partial = nullptr;
        set_error("reduction buffers for %d candidates: %s", cap, cudaGetErrorString(e));
        return OQRL_ERR_NOMEM;
    }
    OQRL_CUDA_CHECK(cudaMemsetAsync(c->red_arrivals, 0, (size_t)cap * sizeof(unsigned), st));
    c->red_capacity = cap;
    c->red_doubles = doubles;
    return OQRL_OK;
}

// ---- spec validation -----------------------------------------------------------
static int check_spec(const oqrl_spec *sp)
{
    if (!sp) { set_error("spec is NULL"); return OQRL_ERR_INVALID; }
    if (sp->n_qubits < 1 || sp->n_qubits > 30) { set_error("n_qubits=%d outside 1..30", sp->n_qubits); return OQRL_ERR_INVALID; }
    if (sp->n_layers < 0) { set_error("n_layers=%d < 0", sp->n_layers); return OQRL_ERR_INVALID; }
    if (sp->n_out < 1 || sp->n_out > sp->n_qubits || sp->n_out > kMaxOut) {
        set_error("n_out=%d outside 1..min(n_qubits,%d)", sp->n_out, kMaxOut); return OQRL_ERR_INVALID;
    }
    if (sp->n_feat < 0) { set_error("n_feat=%d < 0", sp->n_feat); return OQRL_ERR_INVALID; }
    if (sp->feature_map != OQRL_FEAT_FIRST_K && sp->feature_map != OQRL_FEAT_TILED) {
        set_error("feature_map=%d unknown", sp->feature_map); return OQRL_ERR_INVALID;
    }
    if (sp->feature_map == OQRL_FEAT_FIRST_K && sp->n_feat > sp->n_qubits) {
        // qml.AngleEmbedding raises when features outnumber wires.
        set_error("n_feat=%d exceeds n_qubits=%d (AngleEmbedding semantics)", sp->n_feat, sp->n_qubits);
        return OQRL_ERR_INVALID;
    }
    if (sp->entangler != OQRL_ENT_SEL_CNOT && sp->entangler != OQRL_ENT_CZ_RING) {
        set_error("entangler=%d unknown", sp->entangler); return OQRL_ERR_INVALID;
    }
    if (sp->reupload != 0 && sp->reupload != 1) { set_error("reupload=%d must be 0 or 1", sp->reupload); return OQRL_ERR_INVALID; }
    if (sp->reserved & ~7) { set_error("reserved must be 0 (measurement-only switches: bit 0 disables last-layer pruning, bit 1 selects the general shared-memory kernel, bit 2 the direct-load register kernel)"); return OQRL_ERR_INVALID; }
    return OQRL_OK;
}

static int kernel_class(const oqrl_spec &sp) { return sp.n_qubits <= 4 ? 0 : (sp.n_qubits <= 13 ? 1 : 2); }
// class 1 has two implementations; spec.reserved bit 1 forces the general one (pqc_smem.cu) for A/B runs.
static bool use_blk(const oqrl_spec &sp) { return !(sp.reserved & 2) && blk_supports(sp); }

// ---- small shared kernels --------------------------------------------------------
__global__ void trig_table_kernel(const float *__restrict__ obs, const float *__restrict__ next_obs,
                                  const int32_t *__restrict__ act, const float *__restrict__ rew,
                                  const float *__restrict__ term, long long n_rows, int n_feat,
                                  float4 *__restrict__ trig, float4 *__restrict__ meta)
{
    const long long total = n_rows * n_feat;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float c0, s0, c1, s1;
        sincosf(0.5f * obs[i], &s0, &c0);
        sincosf(0.5f * next_obs[i], &s1, &c1);
        trig[i] = make_float4(c0, c1, s0, s1);
    }
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_rows; j += (long long)gridDim.x * blockDim.x)
        meta[j] = make_float4(rew[j], term[j], __int_as_float(act[j]), 0.f);
}

int launch_trig_table(const float *d_obs, const float *d_next_obs, const int32_t *d_act, const float *d_rew,
                      const float *d_term, int64_t n_rows, int n_feat, float4 *d_trig, float4 *d_meta, cudaStream_t st)
{
    if (n_rows <= 0) return OQRL_OK;
    long long total = (long long)n_rows * (n_feat > 0 ? n_feat : 1);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    trig_table_kernel<<<blocks, 256, 0, st>>>(d_obs, d_next_obs, d_act, d_rew, d_term, (long long)n_rows, n_feat, d_trig, d_meta);
    OQRL_LAUNCH_CHECK("trig_table_kernel");
    return OQRL_OK;
}

__global__ void td_finalize_kernel(const double *__restrict__ partial, int n_cand, int n_chunks, int n_trans,
                                   float *__restrict__ fitness)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_cand) return;
    double tot = 0.0;
    for (int c = 0; c < n_chunks; ++c) tot += partial[(size_t)p * n_chunks + c];
    fitness[p] = (float)(-tot / (double)n_trans);
}

int launch_td_finalize(const double *d_partial, int n_cand, int n_chunks, int n_trans, float *d_fitness, cudaStream_t st)
{
    td_finalize_kernel<<<(n_cand + 127) / 128, 128, 0, st>>>(d_partial, n_cand, n_chunks, n_trans, d_fitness);
    OQRL_LAUNCH_CHECK("td_finalize_kernel");
    return OQRL_OK;
}

__global__ void gather_rows_kernel(const float *__restrict__ obs, const float *__restrict__ next_obs,
                                   const float4 *__restrict__ meta, const int32_t *__restrict__ idx, int n_idx,
                                   int n_feat, long long n_rows, float *o_obs, float *o_next, int32_t *o_act,
                                   float *o_rew, float *o_term)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_idx; j += gridDim.x * blockDim.x) {
        const long long row = idx[j];
        if (row < 0 || row >= n_rows) continue;  // rejected on the host for host indices; defensive here
        for (int f = 0; f < n_feat; ++f) {
            if (o_obs) o_obs[(size_t)j * n_feat + f] = obs[row * n_feat + f];
            if (o_next) o_next[(size_t)j * n_feat + f] = next_obs[row * n_feat + f];
        }
        const float4 m = meta[row];
        if (o_rew) o_rew[j] = m.x;
        if (o_term) o_term[j] = m.y;
        if (o_act) o_act[j] = __float_as_int(m.z);
    }
}

__global__ void gather_meta_kernel(const float *__restrict__ obs, const float *__restrict__ next_obs,
                                   const float4 *__restrict__ meta, const int32_t *__restrict__ idx, int n_idx, int n_feat,
                                   long long n_rows, float *o_obs, float *o_next, float4 *o_meta)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_idx; j += gridDim.x * blockDim.x) {
        long long row = idx ? idx[j] : j;
        const bool bad = row < 0 || row >= n_rows;   // same policy as checked_row(): read row 0, poison the result
        if (bad) row = 0;
        for (int f = 0; f < n_feat; ++f) {
            o_obs[(size_t)j * n_feat + f] = obs[row * n_feat + f];
            o_next[(size_t)j * n_feat + f] = next_obs[row * n_feat + f];
        }
        float4 m = meta[row];
        if (bad) m.x = __int_as_float(0x7fc00000);
        o_meta[j] = m;
    }
}

// Next GA population in one launch: elites copied, children = crossover of two ranked parents + mutation noise.
__global__ void ga_breed_kernel(const float *__restrict__ pop, int n_pop, int n_genes, const int32_t *__restrict__ order,
                                int n_elite, const int32_t *__restrict__ pair, int pool, const uint8_t *__restrict__ mask,
                                const double *__restrict__ noise, float *__restrict__ next)
{
    const long long total = (long long)n_pop * n_genes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / n_genes), g = (int)(i - (long long)row * n_genes);
        if (row < n_elite) {
            next[i] = pop[(size_t)order[row] * n_genes + g];
        } else {
            const int c = row - n_elite, j = c >> 1, second = c & 1;
            // parent ranks are host draws: a rank outside the pool reads rank 0 instead of past `order`
            const int ra = pair[2 * j], rb = pair[2 * j + 1];
            const float a = pop[(size_t)order[(unsigned)ra < (unsigned)pool ? ra : 0] * n_genes + g];
            const float b = pop[(size_t)order[(unsigned)rb < (unsigned)pool ? rb : 0] * n_genes + g];
            const bool m = mask[(size_t)j * n_genes + g] != 0;
            const float base = (m != (second != 0)) ? a : b;          // child1 = where(m, a, b), child2 = where(m, b, a)
            // optim.py:193-197: float32 tensor.add_(float64 noise) -> computed in float64, stored as float32
            next[i] = (float)((double)base + noise[((size_t)j * 2 + second) * n_genes + g]);
        }
    }
}

// Selection on the device (optim.py:207 `np.argsort(fitness)[::-1]`, of which the GA uses the first few entries):
// order[r] = index of the candidate with the r-th largest fitness, r < k.  Rank by counting, in the order a STABLE
// ascending sort read backwards gives (numpy's default sort leaves exact ties unspecified and machine-dependent;
// this is the reproducible member of the orders it may return): exact ties resolve to the HIGHER index, NaN
// sorts above +inf (np.argsort puts NaN last, [::-1] makes it first: a NaN candidate becomes the elite, as in the
// reference).  Keys are compared as monotone integers so that NaN and ties need no special cases in the loop.
__device__ __forceinline__ unsigned rank_key(float f)
{
    if (f != f) return 0xffffffffu;                       // every NaN: one key above +inf
    if (f == 0.f) f = 0.f;                                // -0.0 == +0.0 in numpy's comparison
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void ga_rank_kernel(const float *__restrict__ fit, int n, int k, int32_t *__restrict__ order)
{
    __shared__ unsigned tile[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned ki = i < n ? rank_key(fit[i]) : 0u;
    int rank = 0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < n ? rank_key(fit[j]) : 0u;
        __syncthreads();
        const int lim = min(256, n - j0);
        for (int t = 0; t < lim; ++t) {
            const unsigned v = tile[t];
            rank += (v > ki) || (v == ki && j0 + t > i);
        }
        __syncthreads();
    }
    if (i < n && rank < k) order[rank] = i;
}

// Peer-mapped flag vectors of the cross-rank barrier (see oqrl_peer_barrier).
struct PeerFlags {
    unsigned *ptr[kMaxPeers];
};
// Thread p tells peer p "rank `rank` has finished epoch `epoch`" (release store into peer p's flag vector) and waits
// until peer p has told this rank the same.  Epochs only grow, so the flags never need a reset.
__device__ __forceinline__ void peer_signal_and_wait(const PeerFlags &flags, int n_peers, int rank, unsigned epoch, unsigned *timed_out)
{
    const int p = threadIdx.x;
    if (p >= n_peers) return;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.ptr[p] + rank), "r"(epoch) : "memory");
    const unsigned *mine = flags.ptr[rank] + p;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > (20LL << 30)) { *timed_out = 1u; break; }   // ~10 s: a peer died; do not hang the GPU
    }
}

// Selection of the k <= 8 best (the GA needs five: two elites, a parent pool of five) in ONE block, whatever the
// population size: every thread keeps the k best of its strided share as (key << 32 | index) composites -- larger key
// first, larger index on ties, the order of rank_key / ga_rank_kernel -- then k rounds of a block-wide maximum pop
// the winners.  With WITH_PEERS the kernel closes a generation of oqrl_fitness_scatter itself: launched with
// programmatic stream serialization it is resident while the fitness kernel's tail still runs, parks in
// griddepcontrol.wait, signals the peers, waits for their flags and only then reads the exchanged vector -- the
// barrier is no launch of its own.
constexpr int kTopK = 8;
template <bool WITH_PEERS>
__global__ void __launch_bounds__(1024) ga_topk_kernel(const float *__restrict__ fit, int n, int k, int32_t *__restrict__ order,
                                                       PeerFlags flags, int n_peers, int rank, unsigned epoch, unsigned *timed_out)
{
    __shared__ unsigned long long warp_best[32];
    __shared__ unsigned long long winner;
    if (WITH_PEERS) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        peer_signal_and_wait(flags, n_peers, rank, epoch, timed_out);
        __syncthreads();
    }
    unsigned long long best[kTopK];      // descending; 0 = empty
#pragma unroll
    for (int j = 0; j < kTopK; ++j) best[j] = 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        // + 1 keeps 0 free as the "empty" marker whatever the key
        unsigned long long c = (((unsigned long long)rank_key(WITH_PEERS ? __ldcg(fit + i) : fit[i]) << 32) | (unsigned)i) + 1ull;
#pragma unroll
        for (int j = 0; j < kTopK; ++j) {
            if (c > best[j]) { const unsigned long long t = best[j]; best[j] = c; c = t; }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < k; ++r) {
        unsigned long long m = best[0];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, off);
            m = o > m ? o : m;
        }
        if (lane == 0) warp_best[warp] = m;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = lane < (int)(blockDim.x >> 5) ? warp_best[lane] : 0ull;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, w, off);
                w = o > w ? o : w;
            }
            if (lane == 0) { winner = w; order[r] = (int32_t)(unsigned)((w - 1ull) & 0xffffffffull); }
        }
        __syncthreads();
        if (best[0] == winner) {         // composites are unique (they carry the index): exactly one thread pops
#pragma unroll
            for (int j = 0; j < kTopK - 1; ++j) best[j] = best[j + 1];
            best[kTopK - 1] = 0ull;
        }
        __syncthreads();
    }
}

// 32-bit mix (splitmix-style finaliser) and a counter-based stream on top of it: value = f(seed, generation, lane
// of randomness, index).  Used only by the vectorised-RNG breeding mode, which has no stream parity with the
// reference by construction (same distributions: optim.py:175-197).
__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t rnd_u32(uint64_t seed, uint32_t gen, uint32_t stream, uint64_t idx)
{
    uint32_t h = mix32((uint32_t)seed ^ 0x9e3779b9u);
    h = mix32(h ^ (uint32_t)(seed >> 32));
    h = mix32(h ^ gen * 0x85ebca6bu);
    h = mix32(h ^ stream * 0xc2b2ae35u);
    h = mix32(h ^ (uint32_t)idx);
    h = mix32(h ^ (uint32_t)(idx >> 32) * 0x27d4eb2fu);
    return h;
}
__device__ __forceinline__ float rnd_unit(uint32_t u) { return ((u >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

// Breeding with the randomness generated in the kernel: two distinct parents out of the pool per pair, uniform
// crossover mask shared by the two children, independent N(0, sigma) mutation per gene (Box-Muller).
__global__ void ga_breed_random_kernel(const float *__restrict__ pop, int n_pop, int n_genes, const int32_t *__restrict__ order,
                                       int n_elite, int pool, float crossover_rate, float sigma, uint64_t seed, uint32_t gen,
                                       float *__restrict__ next)
{
    const long long total = (long long)n_pop * n_genes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / n_genes), g = (int)(i - (long long)row * n_genes);
        if (row < n_elite) {
            next[i] = pop[(size_t)order[row] * n_genes + g];
            continue;
        }
        const int c = row - n_elite, j = c >> 1, second = c & 1;
        int pa = (int)(rnd_unit(rnd_u32(seed, gen, 0u, (uint64_t)j)) * pool);
        int pb = (int)(rnd_unit(rnd_u32(seed, gen, 1u, (uint64_t)j)) * (pool - 1));
        pa = min(pa, pool - 1); pb = min(pb, pool - 2);
        pb += pb >= pa;
        const uint64_t cell = (uint64_t)j * n_genes + g;
        const bool m = rnd_unit(rnd_u32(seed, gen, 2u, cell)) < crossover_rate;
        const float a = pop[(size_t)order[pa] * n_genes + g], b = pop[(size_t)order[pb] * n_genes + g];
        const float base = (m != (second != 0)) ? a : b;
        const float u1 = rnd_unit(rnd_u32(seed, gen, 3u, cell * 2 + second)), u2 = rnd_unit(rnd_u32(seed, gen, 4u, cell * 2 + second));
        const float z = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
        next[i] = (float)((double)base + (double)(sigma * z));
    }
}

// best individual of a generation: row d_order[0] of the population
__global__ void ga_copy_best_kernel(const float *__restrict__ pop, int n_genes, const int32_t *__restrict__ order, float *__restrict__ best)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_genes) best[g] = pop[(size_t)order[0] * n_genes + g];
}

// Executed FP32 flop per circuit evaluation of the kernel `sp` dispatches to, and HBM state sweeps per circuit for
// the HBM-resident class.  Internal: exported to bench.py / tests by the probe library (csrc/probe.cu).
int work_model(const oqrl_spec *sp, double *flops_per_eval, double *state_sweeps)
{
    int rc = check_spec(sp);
    if (rc) return rc;
    double f = 0, s = 0;
    switch (kernel_class(*sp)) {
    case 0: f = reg_flops_per_eval(*sp); break;
    case 1: f = use_blk(*sp) ? blk_flops_per_eval(*sp) : smem_flops_per_eval(*sp); break;
    default: f = hbm_flops_per_eval(*sp); s = hbm_sweeps_per_eval(*sp); break;
    }
    if (flops_per_eval) *flops_per_eval = f;
    if (state_sweeps) *state_sweeps = s;
    return OQRL_OK;
}
int device_sm_count(int *sm_count)
{
    DeviceCtx *c;
    int rc = current_ctx(&c);
    if (rc) return rc;
    *sm_count = c->sm_count;
    return OQRL_OK;
}

}  // namespace oqrl

using namespace oqrl;

// ---- dataset handle -----------------------------------------------------------
struct oqrl_dataset {
    int device;
    int64_t n_rows;
    int32_t n_feat;
    float *d_obs, *d_next_obs;
    float4 *d_trig, *d_meta;
    int32_t act_min, act_max;   // range of the stored actions, checked against spec.n_out by the fitness entry points
};

extern "C" {

int oqrl_abi_version(void) { return OQRL_ABI_VERSION; }
const char *oqrl_last_error(void) { return g_err; }
int64_t oqrl_launch_count(void) { return (int64_t)g_launches.load(); }

int oqrl_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor)
{
    DeviceCtx *c;
    int rc = current_ctx(&c);
    if (rc) return rc;
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    return OQRL_OK;
}

int oqrl_dataset_upload(const float *h_obs, const float *h_next_obs, const int32_t *h_act, const float *h_rew,
                        const float *h_term, int64_t n_rows, int32_t n_feat, void *stream, oqrl_dataset **out)
{
    if (!out) { set_error("out is NULL"); return OQRL_ERR_INVALID; }
    *out = nullptr;
    if (!h_obs || !h_next_obs || !h_act || !h_rew || !h_term) { set_error("NULL column"); return OQRL_ERR_INVALID; }
    if (n_rows <= 0 || n_rows > (int64_t)0x7fffffff) { set_error("n_rows=%lld outside 1..2^31-1", (long long)n_rows); return OQRL_ERR_INVALID; }
    if (n_feat < 1 || n_feat > 64) { set_error("n_feat=%d outside 1..64", n_feat); return OQRL_ERR_INVALID; }
    DeviceCtx *c;
    int rc = current_ctx(&c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    oqrl_dataset *ds = new (std::nothrow) oqrl_dataset();
    if (!ds) { set_error("host allocation failed"); return OQRL_ERR_NOMEM; }
    memset(ds, 0, sizeof(*ds));
    ds->device = c->device; ds->n_rows = n_rows; ds->n_feat = n_feat;
    ds->act_min = ds->act_max = h_act[0];
    for (int64_t j = 1; j < n_rows; ++j) {
        if (h_act[j] < ds->act_min) ds->act_min = h_act[j];
        if (h_act[j] > ds->act_max) ds->act_max = h_act[j];
    }
    const size_t fb = (size_t)n_rows * n_feat * sizeof(float);
    int32_t *d_act = nullptr; float *d_rew = nullptr, *d_term = nullptr;
    cudaError_t e = cudaSuccess;
    auto fail = [&](const char *what) {
        set_error("dataset upload: %s failed: %s", what, cudaGetErrorString(e));
        cudaGetLastError();
        cudaFree(d_act); cudaFree(d_rew); cudaFree(d_term);
        oqrl_dataset_free(ds);
        return e == cudaErrorMemoryAllocation ? OQRL_ERR_NOMEM : OQRL_ERR_CUDA;
    };
    if ((e = cudaMalloc(&ds->d_obs, fb)) != cudaSuccess) return fail("cudaMalloc(obs)");
    if ((e = cudaMalloc(&ds->d_next_obs, fb)) != cudaSuccess) return fail("cudaMalloc(next_obs)");
    if ((e = cudaMalloc(&ds->d_trig, (size_t)n_rows * n_feat * sizeof(float4))) != cudaSuccess) return fail("cudaMalloc(trig)");
    if ((e = cudaMalloc(&ds->d_meta, (size_t)n_rows * sizeof(float4))) != cudaSuccess) return fail("cudaMalloc(meta)");
    if ((e = cudaMalloc(&d_act, (size_t)n_rows * sizeof(int32_t))) != cudaSuccess) return fail("cudaMalloc(act)");
    if ((e = cudaMalloc(&d_rew, (size_t)n_rows * sizeof(float))) != cudaSuccess) return fail("cudaMalloc(rew)");
    if ((e = cudaMalloc(&d_term, (size_t)n_rows * sizeof(float))) != cudaSuccess) return fail("cudaMalloc(term)");
    if ((e = cudaMemcpyAsync(ds->d_obs, h_obs, fb, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D obs");
    if ((e = cudaMemcpyAsync(ds->d_next_obs, h_next_obs, fb, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D next_obs");
    if ((e = cudaMemcpyAsync(d_act, h_act, (size_t)n_rows * sizeof(int32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D act");
    if ((e = cudaMemcpyAsync(d_rew, h_rew, (size_t)n_rows * sizeof(float), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D rew");
    if ((e = cudaMemcpyAsync(d_term, h_term, (size_t)n_rows * sizeof(float), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D term");
    rc = launch_trig_table(ds->d_obs, ds->d_next_obs, d_act, d_rew, d_term, n_rows, n_feat, ds->d_trig, ds->d_meta, st);
    if (rc) { cudaFree(d_act); cudaFree(d_rew); cudaFree(d_term); oqrl_dataset_free(ds); return rc; }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail("synchronize");
    cudaFree(d_act); cudaFree(d_rew); cudaFree(d_term);
    *out = ds;
    return OQRL_OK;
}

int oqrl_dataset_free(oqrl_dataset *ds)
{
    if (!ds) return OQRL_OK;
    cudaFree(ds->d_obs); cudaFree(ds->d_next_obs); cudaFree(ds->d_trig); cudaFree(ds->d_meta);
    cudaGetLastError();
    delete ds;
    return OQRL_OK;
}

int oqrl_dataset_shape(const oqrl_dataset *ds, int64_t *n_rows, int32_t *n_feat)
{
    if (!ds) { set_error("dataset is NULL"); return OQRL_ERR_INVALID; }
    if (n_rows) *n_rows = ds->n_rows;
    if (n_feat) *n_feat = ds->n_feat;
    return OQRL_OK;
}

int oqrl_dataset_gather(const oqrl_dataset *ds, const int32_t *d_idx, int32_t n_idx, float *d_obs, float *d_next_obs,
                        int32_t *d_act, float *d_rew, float *d_term, void *stream)
{
    if (!ds || !d_idx) { set_error("dataset or idx is NULL"); return OQRL_ERR_INVALID; }
    if (n_idx < 0) { set_error("n_idx < 0"); return OQRL_ERR_INVALID; }
    if (n_idx == 0) return OQRL_OK;
    gather_rows_kernel<<<(n_idx + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        ds->d_obs, ds->d_next_obs, ds->d_meta, d_idx, n_idx, ds->n_feat, (long long)ds->n_rows, d_obs, d_next_obs, d_act,
        d_rew, d_term);
    OQRL_LAUNCH_CHECK("gather_rows_kernel");
    return OQRL_OK;
}

int oqrl_kernel_class(const oqrl_spec *spec)
{
    int rc = check_spec(spec);
    return rc ? rc : kernel_class(*spec);
}

const char *oqrl_kernel_name(const oqrl_spec *spec)
{
    if (check_spec(spec)) return nullptr;
    switch (kernel_class(*spec)) {
    case 0: return "reg_fitness_kernel";
    case 1: return use_blk(*spec) ? "blk_fitness_kernel" : "smem_fitness_kernel";
    default: return "hbm_pass_kernel";
    }
}

int oqrl_qvalues(const oqrl_spec *spec, const float *d_params, int32_t n_cand, const float *d_x, int32_t n_rows,
                 float *d_q, void *stream)
{
    int rc = check_spec(spec);
    if (rc) return rc;
    if (n_cand < 0 || n_rows < 0) { set_error("negative count"); return OQRL_ERR_INVALID; }
    if (n_cand == 0 || n_rows == 0) return OQRL_OK;  // empty batch: nothing to write
    if (!d_params && spec->n_layers > 0) { set_error("params is NULL"); return OQRL_ERR_INVALID; }
    if (!d_q || (!d_x && spec->n_feat > 0)) { set_error("x or q is NULL"); return OQRL_ERR_INVALID; }
    if (n_cand > 65535 * 64) { set_error("n_cand too large"); return OQRL_ERR_INVALID; }
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (kernel_class(*spec)) {
    case 0: return reg_qvalues(*spec, d_params, n_cand, d_x, n_rows, d_q, st);
    case 1:
        if (use_blk(*spec)) return blk_qvalues(*spec, d_params, n_cand, d_x, n_rows, d_q, st);
        return smem_qvalues(*spec, d_params, n_cand, d_x, n_rows, d_q, st);
    default: {
        size_t bytes = hbm_workspace_bytes(*spec, n_cand * n_rows);
        void *ws;
        if ((rc = ensure_workspace(c, bytes, st, &ws))) return rc;
        return hbm_qvalues(*spec, d_params, n_cand, d_x, n_rows, d_q, nullptr, ws, c->workspace_bytes, c->sm_count, st);
    }
    }
}

int oqrl_statevector(const oqrl_spec *spec, const float *d_params, const float *d_x, float *d_state, void *stream)
{
    int rc = check_spec(spec);
    if (rc) return rc;
    if (!d_state) { set_error("state is NULL"); return OQRL_ERR_INVALID; }
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = hbm_workspace_bytes(*spec, 1);
    void *ws;
    if ((rc = ensure_workspace(c, bytes, st, &ws))) return rc;
    return hbm_qvalues(*spec, d_params, 1, d_x, 1, nullptr, d_state, ws, c->workspace_bytes, c->sm_count, st);
}

__global__ void scatter_to_peers_kernel(const float *__restrict__ local, int n_cand, FitnessOut out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cand) out.store(i, local[i]);
}

static int fitness_dispatch(DeviceCtx *c, const oqrl_spec &sp, const float *d_params, int n_cand, const TransitionView &tv,
                            float gamma, float *d_fitness, double *d_partial, cudaStream_t st, const FitnessOut *peers = nullptr)
{
    switch (kernel_class(sp)) {
    case 0: {
        // chunked only while the candidates alone cannot fill ~8 waves of resident warps
        const bool chunked = n_cand < 8 * 16 * c->sm_count;
        if (chunked) {
            int rc = ensure_reduce_buffers(c, n_cand, (tv.n_trans + 255) / 256, st);
            if (rc) return rc;
        }
        FitnessOut out;
        if (peers) out = *peers;
        else { out.ptr[0] = d_fitness; out.n = 1; out.offset = 0; }
        return reg_fitness(sp, d_params, n_cand, tv, gamma, out, chunked ? c->red_partial : nullptr,
                           chunked ? c->red_arrivals : nullptr, c->red_doubles, c->sm_count, st);
    }
    case 1:
        if (use_blk(sp)) return blk_fitness(sp, d_params, n_cand, tv, gamma, d_fitness, d_partial, c->sm_count, st);
        return smem_fitness(sp, d_params, n_cand, tv, gamma, d_fitness, d_partial, c->sm_count, st);
    default:
        set_error("internal: HBM-resident class is dispatched by hbm_fitness");
        return OQRL_ERR_INVALID;
    }
}

// TD fitness from precomputed Q-values: q, qn [P][T][n_out]; one CTA per candidate, fixed-order sum.
__global__ void td_from_q_kernel(const float *__restrict__ q, const float *__restrict__ qn, const float4 *__restrict__ meta,
                                 int n_trans, int n_out, float gamma, float *__restrict__ fitness)
{
    __shared__ double part[8];
    const int cand = blockIdx.x;
    double acc = 0.0;
    for (int t = threadIdx.x; t < n_trans; t += blockDim.x) {
        const float4 m = meta[t];
        const float *qa = q + ((size_t)cand * n_trans + t) * n_out;
        const float *qb = qn + ((size_t)cand * n_trans + t) * n_out;
        float mx = -INFINITY;
        for (int i = 0; i < n_out; ++i) mx = fmaxf(mx, qb[i]);
        const float d = qa[__float_as_int(m.z)] - (m.x + (1.f - m.y) * gamma * mx);
        acc += (double)(d * d);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
        fitness[cand] = (float)(-tot / (double)n_trans);
    }
}

// HBM-resident class: Q(s) and Q(s') for every (candidate, transition) through the tile-pass kernels, then TD.
static int hbm_fitness(DeviceCtx *c, const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_obs,
                       const float *d_next_obs, const float4 *d_meta, int n_trans, float gamma, float *d_fitness,
                       cudaStream_t st)
{
    const size_t qn = (size_t)n_cand * n_trans * sp.n_out;
    float *d_q = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_q, 2 * qn * sizeof(float), st);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("Q-value buffer: %s", cudaGetErrorString(e)); return OQRL_ERR_NOMEM; }
    void *ws;
    int rc = ensure_workspace(c, hbm_workspace_bytes(sp, n_cand * n_trans), st, &ws);
    if (!rc) rc = hbm_qvalues(sp, d_params, n_cand, d_obs, n_trans, d_q, nullptr, ws, c->workspace_bytes, c->sm_count, st);
    if (!rc) rc = hbm_qvalues(sp, d_params, n_cand, d_next_obs, n_trans, d_q + qn, nullptr, ws, c->workspace_bytes, c->sm_count, st);
    if (!rc) {
        td_from_q_kernel<<<n_cand, 256, 0, st>>>(d_q, d_q + qn, d_meta, n_trans, sp.n_out, gamma, d_fitness);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("launch of td_from_q_kernel failed"); rc = OQRL_ERR_CUDA; }
    }
    cudaFreeAsync(d_q, st);
    return rc;
}

static int check_fitness_args(const oqrl_spec *spec, const float *d_params, int n_cand, int n_trans, float *d_fitness)
{
    int rc = check_spec(spec);
    if (rc) return rc;
    if (n_cand < 0) { set_error("n_cand < 0"); return OQRL_ERR_INVALID; }
    if (n_trans <= 0) {
        // torch.nn.functional.mse_loss of an empty batch is NaN in the reference; refuse instead.
        set_error("n_trans=%d: fitness of an empty batch is undefined", n_trans);
        return OQRL_ERR_INVALID;
    }
    if (n_cand > 0 && ((!d_params && spec->n_layers > 0) || !d_fitness)) { set_error("params or fitness is NULL"); return OQRL_ERR_INVALID; }
    if (n_cand > 0x7fffffff / 4) { set_error("n_cand too large"); return OQRL_ERR_INVALID; }
    return OQRL_OK;
}

// A resident dataset must fit the circuit: feature count, and every stored action must address a read-out
// (torch.gather raises in the reference, optim.py:158; every kernel family rejects the same way here).
static int check_dataset(const oqrl_spec *spec, const oqrl_dataset *ds)
{
    if (ds->n_feat != spec->n_feat) { set_error("dataset n_feat=%d != spec n_feat=%d", ds->n_feat, spec->n_feat); return OQRL_ERR_INVALID; }
    if (ds->act_min < 0 || ds->act_max >= spec->n_out) {
        set_error("dataset actions span %d..%d but the circuit reads out %d value(s)", ds->act_min, ds->act_max, spec->n_out);
        return OQRL_ERR_INVALID;
    }
    return OQRL_OK;
}

// partial-sum scratch for the transition-split path: [n_cand][<= 4*SMs] doubles
static size_t partial_bytes(const DeviceCtx *c, int n_cand) { return (size_t)n_cand * (size_t)(4 * c->sm_count) * sizeof(double); }

int oqrl_fitness(const oqrl_spec *spec, const float *d_params, int32_t n_cand, const oqrl_dataset *ds,
                 const int32_t *d_idx, int32_t n_trans, float gamma, float *d_fitness, void *stream)
{
    int rc = check_fitness_args(spec, d_params, n_cand, n_trans, d_fitness);
    if (rc) return rc;
    if (!ds) { set_error("dataset is NULL"); return OQRL_ERR_INVALID; }
    if ((rc = check_dataset(spec, ds))) return rc;
    if (!d_idx && n_trans > ds->n_rows) { set_error("n_trans exceeds dataset rows"); return OQRL_ERR_INVALID; }
    if (n_cand == 0) return OQRL_OK;
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    if (c->device != ds->device) { set_error("dataset lives on device %d, current device is %d", ds->device, c->device); return OQRL_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    double *d_partial = nullptr;
    if (n_cand < 4 * c->sm_count) {
        void *ws;
        if ((rc = ensure_workspace(c, partial_bytes(c, n_cand), st, &ws))) return rc;
        d_partial = (double *)ws;
    }
    if (kernel_class(*spec) == 2) {
        // gather the generation's rows once (raw observations + TD metadata), then the tile-pass path
        const size_t ob = ((size_t)n_trans * spec->n_feat * sizeof(float) + 255) & ~(size_t)255;
        const size_t mb = ((size_t)n_trans * sizeof(float4) + 255) & ~(size_t)255;
        char *buf = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&buf, 2 * ob + mb, st);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("gather buffer: %s", cudaGetErrorString(e)); return OQRL_ERR_NOMEM; }
        float *g_obs = (float *)buf, *g_next = (float *)(buf + ob);
        float4 *g_meta = (float4 *)(buf + 2 * ob);
        gather_meta_kernel<<<(n_trans + 127) / 128, 128, 0, st>>>(ds->d_obs, ds->d_next_obs, ds->d_meta, d_idx, n_trans, ds->n_feat,
                                                                 (long long)ds->n_rows, g_obs, g_next, g_meta);
        count_launch();
        rc = hbm_fitness(c, *spec, d_params, n_cand, g_obs, g_next, g_meta, n_trans, gamma, d_fitness, st);
        cudaFreeAsync(buf, st);
        return rc;
    }
    TransitionView tv{ds->d_trig, ds->d_meta, d_idx, n_trans, (int32_t)ds->n_rows};
    return fitness_dispatch(c, *spec, d_params, n_cand, tv, gamma, d_fitness, d_partial, st);
}

int oqrl_fitness_scatter(const oqrl_spec *spec, const float *d_params, int32_t n_cand, const oqrl_dataset *ds,
                         const int32_t *d_idx, int32_t n_trans, float gamma, float *const *d_fitness_peers,
                         int32_t n_peers, int32_t cand_offset, void *stream)
{
    if (!d_fitness_peers || n_peers < 1 || n_peers > kMaxPeers) { set_error("n_peers outside 1..%d", kMaxPeers); return OQRL_ERR_INVALID; }
    if (cand_offset < 0) { set_error("cand_offset < 0"); return OQRL_ERR_INVALID; }
    for (int r = 0; r < n_peers; ++r)
        if (!d_fitness_peers[r]) { set_error("peer buffer %d is NULL", r); return OQRL_ERR_INVALID; }
    int rc = check_fitness_args(spec, d_params, n_cand, n_trans, d_fitness_peers[0]);
    if (rc) return rc;
    if (!ds) { set_error("dataset is NULL"); return OQRL_ERR_INVALID; }
    if ((rc = check_dataset(spec, ds))) return rc;
    if (n_cand == 0) return OQRL_OK;
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    FitnessOut out;
    out.n = n_peers; out.offset = cand_offset;
    for (int r = 0; r < n_peers; ++r) out.ptr[r] = d_fitness_peers[r];
    if (kernel_class(*spec) == 0) {
        // fused: the fitness kernel's last-arriving warp of every candidate stores to all peers
        TransitionView tv{ds->d_trig, ds->d_meta, d_idx, n_trans, (int32_t)ds->n_rows};
        return fitness_dispatch(c, *spec, d_params, n_cand, tv, gamma, nullptr, nullptr, st, &out);
    }
    // other families: local evaluation, then one small peer-store kernel
    float *local = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&local, (size_t)n_cand * sizeof(float), st);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("local fitness buffer: %s", cudaGetErrorString(e)); return OQRL_ERR_NOMEM; }
    rc = oqrl_fitness(spec, d_params, n_cand, ds, d_idx, n_trans, gamma, local, stream);
    if (!rc) {
        scatter_to_peers_kernel<<<(n_cand + 127) / 128, 128, 0, st>>>(local, n_cand, out);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("launch of scatter_to_peers_kernel failed"); rc = OQRL_ERR_CUDA; }
    }
    cudaFreeAsync(local, st);
    return rc;
}

// Cross-rank barrier that closes a generation of oqrl_fitness_scatter.  Thread p tells peer p "rank `rank` has
// finished epoch `epoch`" (release store into peer p's flag vector) and waits until peer p has told this rank the
// same.  Launched with programmatic stream serialization: the block becomes resident while the fitness kernel's
// last CTAs are still running and parks in griddepcontrol.wait, which returns once that kernel has completed and
// its memory operations (the peer stores) are visible -- so the launch latency of the barrier is off the critical
// path.  Epochs only grow, so the flags never need a reset; a rank that is ahead simply leaves a larger value.
__global__ void peer_barrier_kernel(PeerFlags flags, int n_peers, int rank, unsigned epoch, unsigned *timed_out)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    peer_signal_and_wait(flags, n_peers, rank, epoch, timed_out);
}

int oqrl_peer_barrier(uint32_t *const *d_flag_peers, int32_t n_peers, int32_t rank, uint32_t epoch, void *stream)
{
    if (!d_flag_peers || n_peers < 1 || n_peers > kMaxPeers || rank < 0 || rank >= n_peers) { set_error("bad peer barrier arguments"); return OQRL_ERR_INVALID; }
    PeerFlags f;
    for (int r = 0; r < n_peers; ++r) {
        if (!d_flag_peers[r]) { set_error("peer flag buffer %d is NULL", r); return OQRL_ERR_INVALID; }
        f.ptr[r] = d_flag_peers[r];
    }
    // slot n_peers of this rank's own flag vector records a timeout (checked by the host mirror)
    unsigned *timed_out = f.ptr[rank] + n_peers;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, peer_barrier_kernel, f, (int)n_peers, (int)rank, (unsigned)epoch, timed_out);
    count_launch();
    if (e != cudaSuccess) { cudaGetLastError(); set_error("launch of peer_barrier_kernel failed: %s", cudaGetErrorString(e)); return OQRL_ERR_CUDA; }
    return OQRL_OK;
}

int oqrl_fitness_raw(const oqrl_spec *spec, const float *d_params, int32_t n_cand, const float *d_obs,
                     const float *d_next_obs, const int32_t *d_act, const float *d_rew, const float *d_term,
                     int32_t n_trans, float gamma, float *d_fitness, void *stream)
{
    int rc = check_fitness_args(spec, d_params, n_cand, n_trans, d_fitness);
    if (rc) return rc;
    if ((spec->n_feat > 0 && (!d_obs || !d_next_obs)) || !d_act || !d_rew || !d_term) { set_error("NULL transition column"); return OQRL_ERR_INVALID; }
    if (n_cand == 0) return OQRL_OK;
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nf = spec->n_feat > 0 ? spec->n_feat : 1;
    const size_t trig_b = ((size_t)n_trans * nf * sizeof(float4) + 255) & ~(size_t)255;
    const size_t meta_b = ((size_t)n_trans * sizeof(float4) + 255) & ~(size_t)255;
    const size_t part_b = n_cand < 4 * c->sm_count ? partial_bytes(c, n_cand) : 0;
    void *ws;
    if ((rc = ensure_workspace(c, trig_b + meta_b + part_b, st, &ws))) return rc;
    float4 *trig = (float4 *)ws;
    float4 *meta = (float4 *)((char *)ws + trig_b);
    double *d_partial = part_b ? (double *)((char *)ws + trig_b + meta_b) : nullptr;
    if ((rc = launch_trig_table(d_obs, d_next_obs, d_act, d_rew, d_term, n_trans, spec->n_feat, trig, meta, st))) return rc;
    if (kernel_class(*spec) == 2) {
        // the tile-pass path regrows/reuses the shared workspace: keep the metadata in its own buffer
        float4 *m2 = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&m2, (size_t)n_trans * sizeof(float4), st);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("metadata buffer: %s", cudaGetErrorString(e)); return OQRL_ERR_NOMEM; }
        OQRL_CUDA_CHECK(cudaMemcpyAsync(m2, meta, (size_t)n_trans * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        rc = hbm_fitness(c, *spec, d_params, n_cand, d_obs, d_next_obs, m2, n_trans, gamma, d_fitness, st);
        cudaFreeAsync(m2, st);
        return rc;
    }
    TransitionView tv{trig, meta, nullptr, n_trans, n_trans};
    return fitness_dispatch(c, *spec, d_params, n_cand, tv, gamma, d_fitness, d_partial, st);
}

// device-visible address of a pinned + mapped host allocation (UVA), or NULL for anything else
static void *mapped_device_pointer(const void *h)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int oqrl_fitness_host(const oqrl_spec *spec, const float *h_params, int32_t n_cand, const oqrl_dataset *ds,
                      const int32_t *h_idx, int32_t n_trans, float gamma, float *h_fitness, void *stream)
{
    int rc = check_fitness_args(spec, h_params, n_cand, n_trans, h_fitness);
    if (rc) return rc;
    if (!ds || !h_idx) { set_error("dataset or idx is NULL"); return OQRL_ERR_INVALID; }
    for (int t = 0; t < n_trans; ++t)
        if (h_idx[t] < 0 || h_idx[t] >= ds->n_rows) { set_error("idx[%d]=%d outside dataset (%lld rows)", t, h_idx[t], (long long)ds->n_rows); return OQRL_ERR_INVALID; }
    if (n_cand == 0) return OQRL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t D = (size_t)spec->n_layers * spec->n_qubits * 3;
    const size_t pb = ((size_t)n_cand * D * sizeof(float) + 255) & ~(size_t)255;
    const size_t ib = ((size_t)n_trans * sizeof(int32_t) + 255) & ~(size_t)255;
    const size_t fb = ((size_t)n_cand * sizeof(float) + 255) & ~(size_t)255;
    // staging buffers are separate from the shared workspace (which oqrl_fitness may regrow)
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    char *buf = nullptr;
    if ((rc = ensure_staging(c, pb + ib + fb, st, &buf))) return rc;
    cudaError_t e = cudaSuccess;
    float *d_params = (float *)buf; int32_t *d_idx = (int32_t *)(buf + pb); float *d_fit = (float *)(buf + pb + ib);
    rc = OQRL_OK;
    // Pinned, mapped host buffers are used in place: every candidate's parameters are read exactly once per work
    // item (the table build), and its fitness is written once, so the explicit copies and their launch gaps buy
    // nothing; the kernel pulls / pushes them over PCIe while other warps compute.  Pageable memory takes the
    // staged copies.  (The row indices are read by every work item and always go to device memory.)
    static const bool zero_copy = []() { const char *v = getenv("OQRL_HOST_ZEROCOPY"); return !(v && v[0] == '0'); }();
    const float *p_src = nullptr;
    float *f_dst = nullptr;
    if (zero_copy) {
        p_src = (const float *)mapped_device_pointer(h_params);
        f_dst = (float *)mapped_device_pointer(h_fitness);
    }
    if (!p_src && D && (e = cudaMemcpyAsync(d_params, h_params, (size_t)n_cand * D * sizeof(float), cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = OQRL_ERR_CUDA;
    if (!rc && (e = cudaMemcpyAsync(d_idx, h_idx, (size_t)n_trans * sizeof(int32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = OQRL_ERR_CUDA;
    if (!rc) rc = oqrl_fitness(spec, p_src ? p_src : d_params, n_cand, ds, d_idx, n_trans, gamma, f_dst ? f_dst : d_fit, stream);
    if (!rc && !f_dst && (e = cudaMemcpyAsync(h_fitness, d_fit, (size_t)n_cand * sizeof(float), cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = OQRL_ERR_CUDA;
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (rc == OQRL_ERR_CUDA && e != cudaSuccess) set_error("oqrl_fitness_host copy failed: %s", cudaGetErrorString(e));
    if (!rc && e2 != cudaSuccess) { set_error("oqrl_fitness_host synchronize failed: %s", cudaGetErrorString(e2)); rc = OQRL_ERR_CUDA; }
    return rc;
}

// Closes a generation on the rank that wants the exchanged vector on the HOST: the cross-rank flag exchange of
// peer_barrier_kernel, then the whole vector is written into pinned, mapped host memory by the same kernel -- no
// separate barrier launch and no copy-engine hop between the last peer store and the host seeing the result.
__global__ void __launch_bounds__(1024) gather_to_host_kernel(PeerFlags flags, int n_peers, int rank, unsigned epoch, unsigned *timed_out,
                                                              const float *__restrict__ src, float *__restrict__ dst, int n)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    peer_signal_and_wait(flags, n_peers, rank, epoch, timed_out);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcg(src + i);
}

int oqrl_fitness_scatter_host(const oqrl_spec *spec, const float *h_params, int32_t n_cand, const oqrl_dataset *ds,
                              const int32_t *h_idx, int32_t n_trans, float gamma, float *const *d_fitness_peers,
                              int32_t n_peers, int32_t cand_offset, uint32_t *const *d_flag_peers, int32_t rank,
                              uint32_t epoch, float *h_fitness_all, int32_t n_total, void *stream)
{
    if (!d_fitness_peers || n_peers < 1 || n_peers > kMaxPeers || rank < 0 || rank >= n_peers) { set_error("bad peer arguments"); return OQRL_ERR_INVALID; }
    int rc = check_fitness_args(spec, h_params, n_cand, n_trans, d_fitness_peers[0]);
    if (rc) return rc;
    if (!ds || !h_idx) { set_error("dataset or idx is NULL"); return OQRL_ERR_INVALID; }
    if (h_fitness_all && (n_total < cand_offset + n_cand)) { set_error("n_total=%d smaller than this rank's block end %d", n_total, cand_offset + n_cand); return OQRL_ERR_INVALID; }
    for (int t = 0; t < n_trans; ++t)
        if (h_idx[t] < 0 || h_idx[t] >= ds->n_rows) { set_error("idx[%d]=%d outside dataset (%lld rows)", t, h_idx[t], (long long)ds->n_rows); return OQRL_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t D = (size_t)spec->n_layers * spec->n_qubits * 3;
    const size_t pb = ((size_t)n_cand * D * sizeof(float) + 255) & ~(size_t)255;
    const size_t ib = ((size_t)n_trans * sizeof(int32_t) + 255) & ~(size_t)255;
    DeviceCtx *c;
    if ((rc = current_ctx(&c))) return rc;
    char *buf = nullptr;
    if ((rc = ensure_staging(c, pb + ib, st, &buf))) return rc;
    float *d_params = (float *)buf;
    int32_t *d_idx = (int32_t *)(buf + pb);
    cudaError_t e = cudaSuccess;
    // pinned + mapped parameters are read in place by the fitness kernel (once per work item), as in oqrl_fitness_host
    static const bool zero_copy = []() { const char *v = getenv("OQRL_HOST_ZEROCOPY"); return !(v && v[0] == '0'); }();
    const float *p_src = zero_copy ? (const float *)mapped_device_pointer(h_params) : nullptr;
    if (n_cand > 0 && !p_src && D && (e = cudaMemcpyAsync(d_params, h_params, (size_t)n_cand * D * sizeof(float), cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = OQRL_ERR_CUDA;
    if (!rc && (e = cudaMemcpyAsync(d_idx, h_idx, (size_t)n_trans * sizeof(int32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = OQRL_ERR_CUDA;
    if (!rc) rc = oqrl_fitness_scatter(spec, p_src ? p_src : d_params, n_cand, ds, d_idx, n_trans, gamma, d_fitness_peers, n_peers, cand_offset, stream);
    // only the rank that asks for it reads the exchanged vector back (the rank that runs selection on the host)
    float *f_dst = (h_fitness_all && zero_copy) ? (float *)mapped_device_pointer(h_fitness_all) : nullptr;
    if (!rc && f_dst) {
        // pinned + mapped result buffer: barrier and read-back are ONE kernel, launched under the fitness kernel's tail
        if (!d_flag_peers) { set_error("flag buffers are NULL"); return OQRL_ERR_INVALID; }
        PeerFlags f;
        for (int r = 0; r < n_peers; ++r) {
            if (!d_flag_peers[r]) { set_error("peer flag buffer %d is NULL", r); return OQRL_ERR_INVALID; }
            f.ptr[r] = d_flag_peers[r];
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, gather_to_host_kernel, f, (int)n_peers, (int)rank, (unsigned)epoch, f.ptr[rank] + n_peers,
                               (const float *)d_fitness_peers[rank], f_dst, (int)n_total);
        count_launch();
        if (e != cudaSuccess) { cudaGetLastError(); rc = OQRL_ERR_CUDA; }
    } else {
        if (!rc) rc = oqrl_peer_barrier(d_flag_peers, n_peers, rank, epoch, stream);
        if (!rc && h_fitness_all && (e = cudaMemcpyAsync(h_fitness_all, d_fitness_peers[rank], (size_t)n_total * sizeof(float), cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = OQRL_ERR_CUDA;
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (rc == OQRL_ERR_CUDA && e != cudaSuccess) set_error("oqrl_fitness_scatter_host copy failed: %s", cudaGetErrorString(e));
    if (!rc && e2 != cudaSuccess) { set_error("oqrl_fitness_scatter_host synchronize failed: %s", cudaGetErrorString(e2)); rc = OQRL_ERR_CUDA; }
    return rc;
}

int oqrl_ga_breed(const float *d_pop, int32_t n_pop, int32_t n_genes, const int32_t *d_order, int32_t n_order,
                  int32_t n_elite, const int32_t *d_pair, int32_t parent_pool, const uint8_t *d_mask, const double *d_noise,
                  int32_t n_pairs, float *d_next, void *stream)
{
    if (n_order < 1 || n_order > n_pop || n_elite > n_order || parent_pool < 0 || parent_pool > n_order) {
        set_error("d_order holds %d ranks: need n_elite=%d <= n_order, parent_pool=%d <= n_order <= n_pop=%d", n_order, n_elite, parent_pool, n_pop);
        return OQRL_ERR_INVALID;
    }
    if (!d_pop || !d_next || !d_order) { set_error("NULL population / order"); return OQRL_ERR_INVALID; }
    if (d_pop == d_next) { set_error("d_next must not alias d_pop"); return OQRL_ERR_INVALID; }
    if (n_pop <= 0 || n_genes <= 0 || n_elite < 0 || n_elite > n_pop || n_pairs < 0) { set_error("bad GA sizes"); return OQRL_ERR_INVALID; }
    if (n_elite + 2 * (long long)n_pairs < n_pop) { set_error("n_elite + 2*n_pairs = %lld rows cannot fill a population of %d", n_elite + 2LL * n_pairs, n_pop); return OQRL_ERR_INVALID; }
    if (n_pairs > 0 && (!d_pair || !d_mask || !d_noise)) { set_error("NULL pair / mask / noise"); return OQRL_ERR_INVALID; }
    DeviceCtx *c;
    int rc = current_ctx(&c);
    if (rc) return rc;
    const long long total = (long long)n_pop * n_genes;
    int blocks = (int)((total + 255) / 256);
    if (blocks > c->sm_count * 16) blocks = c->sm_count * 16;
    ga_breed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_pop, n_pop, n_genes, d_order, n_elite, d_pair, parent_pool, d_mask, d_noise, d_next);
    OQRL_LAUNCH_CHECK("ga_breed_kernel");
    return OQRL_OK;
}

int oqrl_ga_rank(const float *d_fitness, int32_t n, int32_t k, int32_t *d_order, void *stream)
{
    if (!d_fitness || !d_order || n <= 0 || k <= 0 || k > n) { set_error("bad ga_rank arguments"); return OQRL_ERR_INVALID; }
    if (k <= kTopK) {
        // what the GA asks for (k = 5): one block, O(n) -- the counting kernel below is O(n^2) and only serves large k
        PeerFlags none = {};
        ga_topk_kernel<false><<<1, 1024, 0, (cudaStream_t)stream>>>(d_fitness, n, k, d_order, none, 0, 0, 0u, nullptr);
        OQRL_LAUNCH_CHECK("ga_topk_kernel");
        return OQRL_OK;
    }
    ga_rank_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_fitness, n, k, d_order);
    OQRL_LAUNCH_CHECK("ga_rank_kernel");
    return OQRL_OK;
}

int oqrl_ga_rank_peers(const float *d_fitness, int32_t n, int32_t k, int32_t *d_order, uint32_t *const *d_flag_peers,
                       int32_t n_peers, int32_t rank, uint32_t epoch, void *stream)
{
    if (!d_fitness || !d_order || n <= 0 || k <= 0 || k > n || k > kTopK) { set_error("bad ga_rank_peers arguments (k <= %d)", kTopK); return OQRL_ERR_INVALID; }
    if (!d_flag_peers || n_peers < 1 || n_peers > kMaxPeers || rank < 0 || rank >= n_peers) { set_error("bad peer barrier arguments"); return OQRL_ERR_INVALID; }
    PeerFlags f;
    for (int r = 0; r < n_peers; ++r) {
        if (!d_flag_peers[r]) { set_error("peer flag buffer %d is NULL", r); return OQRL_ERR_INVALID; }
        f.ptr[r] = d_flag_peers[r];
    }
    unsigned *timed_out = f.ptr[rank] + n_peers;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ga_topk_kernel<true>, d_fitness, (int)n, (int)k, d_order, f, (int)n_peers, (int)rank,
                                       (unsigned)epoch, timed_out);
    count_launch();
    if (e != cudaSuccess) { cudaGetLastError(); set_error("launch of ga_topk_kernel failed: %s", cudaGetErrorString(e)); return OQRL_ERR_CUDA; }
    return OQRL_OK;
}

int oqrl_ga_breed_random(const float *d_pop, int32_t n_pop, int32_t n_genes, const int32_t *d_order, int32_t n_elite,
                         int32_t parent_pool, float crossover_rate, float sigma, uint64_t seed, uint32_t generation,
                         float *d_next, void *stream)
{
    if (!d_pop || !d_next || !d_order) { set_error("NULL population / order"); return OQRL_ERR_INVALID; }
    if (d_pop == d_next) { set_error("d_next must not alias d_pop"); return OQRL_ERR_INVALID; }
    if (n_pop <= 0 || n_genes <= 0 || n_elite < 0 || n_elite > n_pop || parent_pool < 2 || parent_pool > n_pop) { set_error("bad GA sizes"); return OQRL_ERR_INVALID; }
    DeviceCtx *c;
    int rc = current_ctx(&c);
    if (rc) return rc;
    const long long total = (long long)n_pop * n_genes;
    int blocks = (int)((total + 255) / 256);
    if (blocks > c->sm_count * 16) blocks = c->sm_count * 16;
    ga_breed_random_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_pop, n_pop, n_genes, d_order, n_elite, parent_pool,
                                                                    crossover_rate, sigma, seed, generation, d_next);
    OQRL_LAUNCH_CHECK("ga_breed_random_kernel");
    return OQRL_OK;
}

int oqrl_ga_run(const oqrl_spec *spec, float *d_pop_a, float *d_pop_b, int32_t n_pop, const oqrl_dataset *ds,
                const int32_t *d_idx, int32_t n_trans, float gamma, int32_t n_generations, int32_t n_elite,
                int32_t parent_pool, const int32_t *d_pair, const uint8_t *d_mask, const double *d_noise, int32_t n_pairs,
                float crossover_rate, float sigma, uint64_t seed, uint32_t generation0, float *d_fitness,
                int32_t *d_order, float *d_best, int32_t *final_in_b, void *stream)
{
    if (!spec || !d_pop_a || !d_pop_b || d_pop_a == d_pop_b || !d_fitness || !d_order || !d_best || !final_in_b) {
        set_error("oqrl_ga_run: NULL or aliased buffers");
        return OQRL_ERR_INVALID;
    }
    if (n_generations <= 0 || n_pop <= 0 || n_elite < 0 || parent_pool < 0) { set_error("oqrl_ga_run: bad sizes"); return OQRL_ERR_INVALID; }
    const bool exact = d_pair || d_mask || d_noise;
    if (exact && (!d_pair || !d_mask || !d_noise)) { set_error("oqrl_ga_run: d_pair, d_mask and d_noise come together"); return OQRL_ERR_INVALID; }
    const int k = n_elite > parent_pool ? n_elite : parent_pool;
    if (k < 1 || k > n_pop) { set_error("oqrl_ga_run: max(n_elite, parent_pool) = %d outside 1..n_pop = %d", k, n_pop); return OQRL_ERR_INVALID; }
    const int n_genes = spec->n_layers * spec->n_qubits * 3;
    if (n_genes <= 0) { set_error("oqrl_ga_run: the circuit has no parameters"); return OQRL_ERR_INVALID; }
    float *cur = d_pop_a, *nxt = d_pop_b;
    for (int g = 0; g < n_generations; ++g) {
        int rc = oqrl_fitness(spec, cur, n_pop, ds, d_idx, n_trans, gamma, d_fitness, stream);
        if (rc) return rc;
        if ((rc = oqrl_ga_rank(d_fitness, n_pop, k, d_order, stream))) return rc;
        if (g == n_generations - 1) {
            ga_copy_best_kernel<<<(n_genes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cur, n_genes, d_order, d_best);
            OQRL_LAUNCH_CHECK("ga_copy_best_kernel");
        }
        if (exact)
            rc = oqrl_ga_breed(cur, n_pop, n_genes, d_order, k, n_elite, d_pair + (size_t)g * n_pairs * 2, parent_pool,
                               d_mask + (size_t)g * n_pairs * n_genes, d_noise + (size_t)g * n_pairs * 2 * n_genes, n_pairs, nxt, stream);
        else
            rc = oqrl_ga_breed_random(cur, n_pop, n_genes, d_order, n_elite, parent_pool, crossover_rate, sigma, seed,
                                      generation0 + (uint32_t)g, nxt, stream);
        if (rc) return rc;
        float *t = cur; cur = nxt; nxt = t;
    }
    *final_in_b = cur == d_pop_b ? 1 : 0;
    return OQRL_OK;
}

}  // extern "C"
