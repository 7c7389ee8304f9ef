// HBM-resident PQC kernels for n >= 14 qubits (BASELINE configs 4 and 5).
//
// The state (complex64, 2^n amplitudes per circuit) lives in HBM.  Work is organised as PASSES:
// a CTA stages one TILE of 2^kT amplitudes in shared memory with fully coalesced 256-byte rows,
// applies every pending gate whose pair direction lies inside the tile (up to kT gates, in
// register blocks of four, exactly like pqc_smem.cu) and writes the tile back -- so a layer
// costs ceil(n / (kT - kB))-ish read+write sweeps of the state instead of n.
//
// Entanglers cost NOTHING here.  A CNOT ring is a GF(2)-linear relabelling of basis states, so
// instead of moving amplitudes the planner tracks the map x = A k between the logical index k
// and the physical address x:
//     gate on logical bit p  ->  pairs (x, x ^ d_p), d_p = A e_p (column p of A);
//     role of x in its pair  ->  logical bit p of k = parity(r_p & x), r_p = row p of A^-1.
// A tile is the affine subspace  base ^ span{e_0..e_{kB-1}, d_p : p in the pass}: it always
// contains the kB low address bits (coalescing) and is closed under every pair direction of the
// pass.  The host planner (build_plan) does the linear algebra once per circuit structure; the
// kernel only XORs precomputed masks.  The CZ ring is a sign folded into the last store of a layer.
//
// The first pass generates the embedded product state (with layer 0's Rot gates folded in, as in
// the other families) instead of reading HBM; the last pass reduces |amp|^2 into <Z_i> instead of
// writing.  Traffic per circuit = (passes - 1) * 2 * 2^n * 8 bytes  (hbm_sweeps_per_eval).
#include <stdlib.h>
#include <string.h>

#include <array>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace oqrl {

namespace {

constexpr int kT = 13;          // log2 amplitudes per tile (64 KiB of complex64 in shared memory)
constexpr int kB = 4;           // log2 amplitudes per contiguous row (128 bytes)
constexpr int kHbmThreads = 256;
constexpr int kMaxGroups = (kT + 3) / 4;
constexpr int kHbmMaxOut = 8;

struct HbmGate {
    int32_t layer, wire;
    uint32_t delta;     // pair direction in tile coordinates (kT bits)
    uint32_t rho;       // role functional in tile coordinates
    uint32_t r_phys;    // role functional on the tile base address (n bits)
    int32_t pad;
};
struct HbmGroup {
    int32_t n_gates;
    int32_t gate[4];
    int32_t n_free;                 // = kT - n_gates
    uint8_t free_pos[kT];           // tile-coordinate bit positions enumerating the register blocks
    uint8_t piv_pos[4];             // the complement: pivot bits of the gate directions, ascending (0xff = unused)
    uint8_t rot_lane[4];            // per gate: lane bit that toggles "start from the partner element" (0xff = none);
                                    // chosen so that the 32 lanes of a warp touch 32 distinct shared-memory banks
    uint8_t pad[3];
};
struct HbmPass {
    int32_t flags;                  // PASS_* bits
    int32_t n_tile_bits;            // n - kT
    uint8_t tile_pos[32];           // physical bit positions enumerating the tiles
    uint32_t row_vec[kT - kB];      // address contribution of each row-index bit (low kB bits are zero)
    int32_t n_gates;
    HbmGate gate[kT];
    int32_t n_groups;
    HbmGroup group[kMaxGroups];
    uint32_t z_mask[kHbmMaxOut];    // final pass: <Z_i> sign = parity(z_mask[i] & x)
    uint32_t a_col[32];             // layout after this pass's layer (for the un-permute test hook)
    uint8_t row_piv[kT - kB];       // pivot bit positions of row_vec, ascending: tile base = tile id with zeros inserted there
    uint8_t pad2[16 - (kT - kB)];
};
enum { PASS_FIRST = 1,     // generate the product state (layer 0 folded in) instead of loading
       PASS_FINAL = 2,     // reduce |amp|^2 into <Z_i> instead of storing
       PASS_CZ = 4,        // CZ-ring sign on the store (this layer's entangler)
       PASS_CZ_PRE = 8 };  // CZ-ring sign right after generation (layer 0's entangler)

inline int parity32(uint32_t v) { return __builtin_popcount(v) & 1; }

// reduced echelon basis over GF(2) with "highest set bit" pivots
struct Echelon {
    std::vector<uint32_t> vec;      // reduced: each pivot bit is set in exactly one vector
    std::vector<int> pivot;
    uint32_t reduce(uint32_t v) const
    {
        for (size_t i = 0; i < vec.size(); ++i)
            if ((v >> pivot[i]) & 1u) v ^= vec[i];
        return v;
    }
    bool add(uint32_t v)
    {
        v = reduce(v);
        if (!v) return false;
        const int p = 31 - __builtin_clz(v);
        for (auto &w : vec)
            if ((w >> p) & 1u) w ^= v;
        vec.push_back(v);
        pivot.push_back(p);
        return true;
    }
    uint32_t pivot_mask() const
    {
        uint32_t m = 0;
        for (int p : pivot) m |= 1u << p;
        return m;
    }
};

unsigned sel_perm_inv(unsigned k, int n, int r)
{
    for (int i = n - 1; i >= 0; --i) {
        int t = (i + r) % n;
        if ((k >> (n - 1 - i)) & 1u) k ^= 1u << (n - 1 - t);
    }
    return k;
}

}  // namespace

static int build_plan_uncached(const oqrl_spec &sp, std::vector<HbmPass> &plan, bool keep_state);

// Host planner: the pass list of one circuit structure (cached: it only depends on the structure).
static int build_plan(const oqrl_spec &sp, std::vector<HbmPass> &plan, bool keep_state = false)
{
    static std::mutex mu;
    static std::map<std::array<int, 5>, std::vector<HbmPass>> cache;
    const std::array<int, 5> key = {sp.n_qubits, sp.n_layers, sp.n_out, sp.entangler, (keep_state ? 1 : 0) | ((sp.reserved & 1) << 1)};
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { plan = it->second; return OQRL_OK; }
    int rc = build_plan_uncached(sp, plan, keep_state);
    if (rc == OQRL_OK) cache[key] = plan;
    return rc;
}

static int build_plan_uncached(const oqrl_spec &sp, std::vector<HbmPass> &plan, bool keep_state)
{
    const int n = sp.n_qubits, L = sp.n_layers;
    if (n < kT + 1 || n > 30) { set_error("HBM family covers %d..30 qubits", kT + 1); return OQRL_ERR_UNSUPPORTED; }
    if (sp.n_out > kHbmMaxOut) { set_error("HBM family supports n_out <= %d", kHbmMaxOut); return OQRL_ERR_UNSUPPORTED; }
    std::vector<uint32_t> a_col(n), ainv_row(n);
    for (int p = 0; p < n; ++p) a_col[p] = ainv_row[p] = 1u << p;
    const uint32_t low_mask = (1u << kB) - 1;
    plan.clear();

    auto make_pass = [&](const std::vector<int> &gate_pos, int layer) {
        HbmPass ps;
        memset(&ps, 0, sizeof(ps));
        // tile subspace: high parts of the gate directions, padded with unit vectors to kT - kB dimensions
        Echelon hi;
        for (int p : gate_pos) hi.add(a_col[p] & ~low_mask);
        for (int pos = n - 1; pos >= kB && (int)hi.vec.size() < kT - kB; --pos) hi.add(1u << pos);
        const uint32_t piv = hi.pivot_mask();
        for (int j = 0; j < kT - kB; ++j) ps.row_vec[j] = hi.vec[j];
        ps.n_tile_bits = 0;
        for (int pos = kB; pos < n; ++pos)
            if (!((piv >> pos) & 1u)) ps.tile_pos[ps.n_tile_bits++] = (uint8_t)pos;
        // gates in tile coordinates
        ps.n_gates = (int)gate_pos.size();
        for (int g = 0; g < ps.n_gates; ++g) {
            const int p = gate_pos[g];
            HbmGate &hg = ps.gate[g];
            hg.layer = layer;
            hg.wire = n - 1 - p;
            uint32_t d = a_col[p], alpha = 0;
            for (int j = 0; j < kT - kB; ++j)
                if ((d >> hi.pivot[j]) & 1u) { alpha |= 1u << j; d ^= hi.vec[j]; }
            // d is now inside the low kB bits
            hg.delta = (alpha << kB) | (d & low_mask);
            uint32_t rho = ainv_row[p] & low_mask;
            for (int j = 0; j < kT - kB; ++j)
                if (parity32(ainv_row[p] & hi.vec[j])) rho |= 1u << (kB + j);
            hg.rho = rho;
            hg.r_phys = ainv_row[p];
        }
        // register blocks of up to four gates, sizes balanced (9 gates -> 3+3+3, not 4+4+1)
        ps.n_groups = 0;
        const int want_groups = (ps.n_gates + 3) / 4;
        for (int g0 = 0, gi = 0; gi < want_groups; ++gi) {
            HbmGroup &gr = ps.group[ps.n_groups++];
            gr.n_gates = ps.n_gates / want_groups + (gi < ps.n_gates % want_groups ? 1 : 0);
            Echelon de;
            for (int i = 0; i < gr.n_gates; ++i) { gr.gate[i] = g0 + i; de.add(ps.gate[g0 + i].delta); }
            g0 += gr.n_gates;
            const uint32_t dp = de.pivot_mask();
            gr.n_free = 0;
            int np = 0;
            memset(gr.piv_pos, 0xff, sizeof(gr.piv_pos));
            for (int pos = 0; pos < kT; ++pos) {
                if (!((dp >> pos) & 1u)) gr.free_pos[gr.n_free++] = (uint8_t)pos;
                else gr.piv_pos[np++] = (uint8_t)pos;
            }
            // Bank-conflict avoidance.  Lane bits 0..4 enumerate the five lowest free positions; if a gate
            // direction has its pivot among the low bits, all lanes would share that bit.  Rotating the block
            // by the gate direction on a lane bit whose free position is >= 5 restores 32 distinct low-5-bit
            // patterns.  Brute force over the (tiny) assignment space, keep the best GF(2) rank.
            memset(gr.rot_lane, 0xff, sizeof(gr.rot_lane));
            {
                const int G = gr.n_gates;
                int best_rank = -1;
                uint8_t best[4] = {0xff, 0xff, 0xff, 0xff};
                int assign[4];
                const int options = 6;   // lane bit 0..4 or "none"
                int total = 1;
                for (int i = 0; i < G; ++i) total *= options;
                for (int code = 0; code < total; ++code) {
                    int cc = code;
                    bool ok = true;
                    uint32_t used = 0;
                    for (int i = 0; i < G; ++i) {
                        assign[i] = cc % options; cc /= options;
                        if (assign[i] < 5) {
                            if ((used >> assign[i]) & 1u) ok = false;
                            used |= 1u << assign[i];
                        }
                    }
                    if (!ok) continue;
                    Echelon e;
                    int rank = 0;
                    for (int j = 0; j < 5 && j < gr.n_free; ++j) {
                        uint32_t col = (1u << gr.free_pos[j]);
                        for (int i = 0; i < G; ++i)
                            if (assign[i] == j) col ^= ps.gate[gr.gate[i]].delta;
                        if (e.add(col & 31u)) ++rank;
                    }
                    if (rank > best_rank) {
                        best_rank = rank;
                        for (int i = 0; i < 4; ++i) best[i] = (i < G && assign[i] < 5) ? (uint8_t)assign[i] : 0xff;
                    }
                    if (rank == 5) break;
                }
                memcpy(gr.rot_lane, best, 4);
            }
        }
        {
            int np = 0;
            for (int pos = kB; pos < n; ++pos)
                if ((piv >> pos) & 1u) ps.row_piv[np++] = (uint8_t)pos;
        }
        plan.push_back(ps);
    };

    bool cz_pre = false;
    for (int l = 0; l < L; ++l) {
        // layer 0's Rot gates are folded into the generated product state: no pass of its own
        if (l > 0) {
            std::vector<int> remaining;
            // last layer: only wires inside the read-out light cone need their gate (exact; see readout_wire_mask)
            const unsigned lm = (l == L - 1 && !keep_state && !(sp.reserved & 1))
                                    ? readout_wire_mask(n, L, sp.n_out, sp.entangler) : 0xffffffffu;
            for (int p = n - 1; p >= 0; --p)
                if ((lm >> (n - 1 - p)) & 1u) remaining.push_back(p);
            while (!remaining.empty()) {
                Echelon hi;
                std::vector<int> take, rest;
                for (int p : remaining) {
                    const uint32_t h = a_col[p] & ~low_mask;
                    if ((int)take.size() < kT && (hi.reduce(h) == 0 || (int)hi.vec.size() < kT - kB)) {
                        hi.add(h);
                        take.push_back(p);
                    } else {
                        rest.push_back(p);
                    }
                }
                // register blocks of exactly four gates are the efficient ones: trim the pass to a multiple of
                // four when the gates pushed back still fit the later passes of this layer (never adds a pass)
                if (!rest.empty() && take.size() > 4 && take.size() % 4 != 0) {
                    const size_t keep = take.size() / 4 * 4;
                    const size_t per_pass = (size_t)(kT - kB);   // what a later pass can always absorb
                    const size_t later = rest.size() + (take.size() - keep);
                    const size_t passes_now = (rest.size() + per_pass - 1) / per_pass;
                    if ((later + per_pass - 1) / per_pass <= passes_now) {
                        for (size_t i = keep; i < take.size(); ++i) rest.insert(rest.begin(), take[i]);
                        take.resize(keep);
                    }
                }
                make_pass(take, l);
                remaining.swap(rest);
            }
        }
        // entangler of layer l
        if (n > 1 && sp.entangler == OQRL_ENT_SEL_CNOT) {
            const int r = sel_range(n, l);
            std::vector<uint32_t> nc(n), nr(n, 0u);
            for (int p = 0; p < n; ++p) {
                uint32_t mi = sel_perm_inv(1u << p, n, r), v = 0;     // A' = A M^-1
                for (int q = 0; q < n; ++q)
                    if ((mi >> q) & 1u) v ^= a_col[q];
                nc[p] = v;
            }
            for (int q = 0; q < n; ++q) {                              // A'^-1 = M A^-1
                const uint32_t mq = sel_perm(1u << q, n, r);
                for (int p = 0; p < n; ++p)
                    if ((mq >> p) & 1u) nr[p] ^= ainv_row[q];
            }
            a_col.swap(nc);
            ainv_row.swap(nr);
        } else if (n > 1 && sp.entangler == OQRL_ENT_CZ_RING) {
            if (l == 0) cz_pre = true;
            else plan.back().flags |= PASS_CZ;
        }
    }
    if (plan.empty()) make_pass({}, L > 0 ? 0 : -1);   // L <= 1: generate and read out in one pass
    if (cz_pre) plan.front().flags |= PASS_CZ_PRE;
    plan.front().flags |= PASS_FIRST;
    HbmPass &fin = plan.back();
    if (!keep_state) fin.flags |= PASS_FINAL;
    for (int i = 0; i < sp.n_out; ++i) fin.z_mask[i] = ainv_row[n - 1 - i];
    for (int p = 0; p < n; ++p) fin.a_col[p] = a_col[p];
    return OQRL_OK;
}

// ---------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------
namespace {

struct HbmArgs {
    const HbmPass *plan;        // device copy of the plan
    int pass;
    float2 *state;              // [n_circ][2^n]
    const float4 *coef;         // [n_circ][max(L,1)][n] SU(2) coefficients {ar, ai, br, bi} (re-upload fused)
    const float4 *vec0;         // [n_circ][n][2]: per-wire 2-vector of the generated product state {re, im, -, -}
    float *zpart;               // [n_circ][n_tiles][n_out] partial <Z_i> sums (final pass)
    int n, n_layers, n_out, n_circ;
};

constexpr int kGenChunkBits = 8;                       // generated state: product of per-8-bit-chunk tables
constexpr int kGenMaxChunks = 4;                       // covers n - kB <= 32
constexpr int kChunks16 = (1 << kT) * 8 / 16;          // 16-byte pieces per tile
constexpr int kPiecesPerThread = kChunks16 / kHbmThreads;
static_assert(kT - kB == 9 && kHbmThreads == 256 && kB == 4, "row/thread mapping below assumes 512 rows x 8 pieces, 256 threads");

// value with a zero bit inserted at every listed position (ascending) -- the coset representative of a
// subspace whose pivot bits are those positions
template <int MAXP>
__device__ __forceinline__ uint32_t insert_zeros(uint32_t v, const uint8_t *piv, int np)
{
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        if (i < np) {
            const int p = piv[i];
            v = ((v >> p) << (p + 1)) | (v & ((1u << p) - 1u));
        }
    }
    return v;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// SU(2) pair update on two register blocks at once (A in the low, B in the high half of every packed value).
// ar, bi are common to both halves (scalar broadcast operands); ai, br carry per-half signs (role swaps).
__device__ __forceinline__ void su2_pair_ab(float ar, float bi, u64 ai, u64 nai, u64 br, u64 nbr,
                                            u64 &x0r, u64 &x0i, u64 &x1r, u64 &x1i)
{
    const u64 sar = bcast2(ar), sbi = bcast2(bi), snbi = bcast2(-bi);
    u64 n0r = fmul2(sar, x0r);
    u64 n0i = fmul2(sar, x0i);
    u64 n1r = fmul2(nbr, x0r);
    u64 n1i = fmul2(nbr, x0i);
    n0r = ffma2(nai, x0i, n0r);
    n0i = ffma2(ai, x0r, n0i);
    n1r = ffma2(snbi, x0i, n1r);
    n1i = ffma2(sbi, x0r, n1i);
    n0r = ffma2(br, x1r, n0r);
    n0i = ffma2(br, x1i, n0i);
    n1r = ffma2(sar, x1r, n1r);
    n1i = ffma2(sar, x1i, n1i);
    n0r = ffma2(snbi, x1i, n0r);
    n0i = ffma2(sbi, x1r, n0i);
    n1r = ffma2(ai, x1i, n1r);
    n1i = ffma2(nai, x1r, n1i);
    x0r = n0r; x0i = n0i; x1r = n1r; x1i = n1i;
}

// Register blocks of 2^G amplitudes: element m of block beta sits at tile coordinate cb(beta) ^ off[m], where
// off[m] is the XOR of the gate directions selected by the bits of m and cb inserts zeros at the pivot bits
// (then rotated per lane, see HbmGroup::rot_lane).  A thread runs TWO blocks at a time (beta and beta + 256)
// packed into fp32x2, out of structure-of-arrays shared-memory planes (32-bit accesses, one bank per lane).
template <int G>
__device__ __forceinline__ void run_group(const HbmPass &ps, const HbmGroup &gr, float *tre, float *tim,
                                          const float4 *s_coef, uint32_t base)
{
    uint32_t off[1 << G];      // element offsets, pre-scaled to bytes (XOR and scaling commute)
    float ar[G], ai[G], br[G], bi[G];
    uint32_t rho[G];
    int rho0[G];
    uint32_t offr = 0;
    off[0] = 0;
    char *pre = reinterpret_cast<char *>(tre), *pim = reinterpret_cast<char *>(tim);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const HbmGate &hg = ps.gate[gr.gate[i]];
#pragma unroll
        for (int m = 0; m < (1 << i); ++m) off[m | (1 << i)] = off[m] ^ (hg.delta << 2);
        rho[i] = hg.rho;
        rho0[i] = __popc(hg.r_phys & base) & 1;
        const float4 c = s_coef[gr.gate[i]];
        ar[i] = c.x; ai[i] = c.y; br[i] = c.z; bi[i] = c.w;
        const int rl = gr.rot_lane[i];
        if (rl != 0xff && ((threadIdx.x >> rl) & 1)) offr ^= hg.delta;
    }
    constexpr int kBlocks = 1 << (kT - G);
    static_assert(kBlocks >= 2 * kHbmThreads, "two blocks per thread per iteration");
#pragma unroll 1
    for (int beta = threadIdx.x; beta < kBlocks; beta += 2 * kHbmThreads) {
        const uint32_t ca = insert_zeros<4>((uint32_t)beta, gr.piv_pos, G) ^ offr;
        const uint32_t cbb = insert_zeros<4>((uint32_t)(beta + kHbmThreads), gr.piv_pos, G) ^ offr;
        const uint32_t ca4 = ca << 2, cb4 = cbb << 2;
        u64 vr[1 << G], vi[1 << G];
#pragma unroll
        for (int m = 0; m < (1 << G); ++m) {
            const uint32_t oa = ca4 ^ off[m], ob = cb4 ^ off[m];
            vr[m] = pack2(*reinterpret_cast<const float *>(pre + oa), *reinterpret_cast<const float *>(pre + ob));
            vi[m] = pack2(*reinterpret_cast<const float *>(pim + oa), *reinterpret_cast<const float *>(pim + ob));
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            // role of element m for gate i is s ^ m_i; s = 1 swaps the roles: X U X = [[conj a, -conj b], [b, a]]
            const bool sa = ((__popc(rho[i] & ca) & 1) ^ rho0[i]) != 0;
            const bool sb = ((__popc(rho[i] & cbb) & 1) ^ rho0[i]) != 0;
            const u64 gai = pack2(sa ? -ai[i] : ai[i], sb ? -ai[i] : ai[i]);
            const u64 gbr = pack2(sa ? -br[i] : br[i], sb ? -br[i] : br[i]);
            const u64 ngai = fneg2(gai), ngbr = fneg2(gbr);
#pragma unroll
            for (int m = 0; m < (1 << G); ++m)
                if (((m >> i) & 1) == 0)
                    su2_pair_ab(ar[i], bi[i], gai, ngai, gbr, ngbr, vr[m], vi[m], vr[m | (1 << i)], vi[m | (1 << i)]);
        }
#pragma unroll
        for (int m = 0; m < (1 << G); ++m) {
            float a0, b0, a1, b1;
            unpack2(vr[m], a0, b0);
            unpack2(vi[m], a1, b1);
            const uint32_t oa = ca4 ^ off[m], ob = cb4 ^ off[m];
            *reinterpret_cast<float *>(pre + oa) = a0; *reinterpret_cast<float *>(pre + ob) = b0;
            *reinterpret_cast<float *>(pim + oa) = a1; *reinterpret_cast<float *>(pim + ob) = b1;
        }
    }
}

__global__ void __launch_bounds__(kHbmThreads, 2)
hbm_pass_kernel(HbmArgs a)
{
    extern __shared__ __align__(16) float tile_planes[];   // re[2^kT] then im[2^kT]
    float *tre = tile_planes, *tim = tile_planes + (1 << kT);
    __shared__ HbmPass ps;
    __shared__ float4 s_coef[kT];
    __shared__ float2 s_vec0[64];                                   // this circuit's per-wire 2-vectors
    __shared__ float2 s_gen[kGenMaxChunks][1 << kGenChunkBits];     // partial products of the generated state
    __shared__ float2 s_low[1 << kB];
    __shared__ float s_z[kHbmMaxOut][kHbmThreads / 32];

    {   // plan of this pass -> shared memory
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.plan + a.pass);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&ps);
        for (int i = threadIdx.x; i < (int)(sizeof(HbmPass) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int n = a.n;
    const int tiles_per_circ = 1 << ps.n_tile_bits;
    const long long total_tiles = (long long)tiles_per_circ * a.n_circ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kHbmThreads / 32;
    const bool first = ps.flags & PASS_FIRST, final_pass = ps.flags & PASS_FINAL, cz = ps.flags & PASS_CZ;
    const bool cz_pre = ps.flags & PASS_CZ_PRE;

    // A thread moves 16-byte pieces (2 amplitudes): piece (row, col16) with row = (tid >> 3) + 32 k, k = 0..15.
    // Row address = XOR of row_vec over the row bits; the five low row bits are fixed per thread.
    const int row_lo = threadIdx.x >> 3, col16 = threadIdx.x & 7;
    uint32_t ra_lo = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j)
        if ((row_lo >> j) & 1) ra_lo ^= ps.row_vec[j];
    const uint32_t rv5 = ps.row_vec[5], rv6 = ps.row_vec[6], rv7 = ps.row_vec[7], rv8 = ps.row_vec[8];
    auto piece_addr = [&](int k) {   // address contribution of piece k (before ^ base)
        uint32_t r = ra_lo;
        if (k & 1) r ^= rv5;
        if (k & 2) r ^= rv6;
        if (k & 4) r ^= rv7;
        if (k & 8) r ^= rv8;
        return r;
    };
    const int n_gen_chunks = (n - kB + kGenChunkBits - 1) / kGenChunkBits;
    int gen_circ = -1;

    for (long long w = blockIdx.x; w < total_tiles; w += gridDim.x) {
        const int circ = (int)(w / tiles_per_circ);
        const uint32_t tile_id = (uint32_t)(w - (long long)circ * tiles_per_circ);
        const uint32_t base = insert_zeros<kT - kB>(tile_id << kB, ps.row_piv, kT - kB);   // zeros at the row pivots
        float2 *st = a.state + ((size_t)circ << n);

        // gate coefficients of this circuit for this pass
        if (threadIdx.x < ps.n_gates) {
            const HbmGate &hg = ps.gate[threadIdx.x];
            s_coef[threadIdx.x] = a.coef[((size_t)circ * max(a.n_layers, 1) + hg.layer) * n + hg.wire];
        }
        if (first && circ != gen_circ) {
            // tables of the generated product state for this circuit (A = identity while the first pass runs)
            const float4 *vec0 = a.vec0 + (size_t)circ * n * 2;
            if (threadIdx.x < 2 * n) { const float4 f = vec0[threadIdx.x]; s_vec0[threadIdx.x] = make_float2(f.x, f.y); }
            __syncthreads();
            if (threadIdx.x < (1 << kB)) {
                float2 p = make_float2(1.f, 0.f);
                for (int pos = 0; pos < kB; ++pos) p = cmul(p, s_vec0[(n - 1 - pos) * 2 + ((threadIdx.x >> pos) & 1)]);
                s_low[threadIdx.x] = p;
            }
            for (int i = threadIdx.x; i < n_gen_chunks << kGenChunkBits; i += kHbmThreads) {
                const int ch = i >> kGenChunkBits, v = i & ((1 << kGenChunkBits) - 1);
                float2 p = make_float2(1.f, 0.f);
                for (int b = 0; b < kGenChunkBits; ++b) {
                    const int pos = kB + ch * kGenChunkBits + b;
                    if (pos < n) p = cmul(p, s_vec0[(n - 1 - pos) * 2 + ((v >> b) & 1)]);
                }
                s_gen[ch][v] = p;
            }
            gen_circ = circ;
        }
        __syncthreads();

        // ---- stage the tile (interleaved complex64 in HBM -> re / im planes in shared memory) ------------------
        if (first) {
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const uint32_t addr = (base ^ piece_addr(k)) + 2 * col16;
                float2 p = make_float2(1.f, 0.f);
                for (int ch = 0; ch < n_gen_chunks; ++ch)
                    p = cmul(p, s_gen[ch][(addr >> (kB + ch * kGenChunkBits)) & ((1 << kGenChunkBits) - 1)]);
                float2 v0 = cmul(p, s_low[2 * col16]), v1 = cmul(p, s_low[2 * col16 + 1]);
                if (cz_pre) {
                    const uint32_t mask = (1u << n) - 1u;
                    const uint32_t k0 = addr, k1 = addr + 1;
                    if (__popc(k0 & (((k0 << 1) | (k0 >> (n - 1))) & mask)) & 1) { v0.x = -v0.x; v0.y = -v0.y; }
                    if (__popc(k1 & (((k1 << 1) | (k1 >> (n - 1))) & mask)) & 1) { v1.x = -v1.x; v1.y = -v1.y; }
                }
                const int e = ((row_lo + 32 * k) << kB) + 2 * col16;
                *reinterpret_cast<float2 *>(tre + e) = make_float2(v0.x, v1.x);
                *reinterpret_cast<float2 *>(tim + e) = make_float2(v0.y, v1.y);
            }
        } else {
            // all 16 pieces of the thread are in flight at once (64 KiB per CTA), coalesced 128-bit loads
            float4 pc[kPiecesPerThread];
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k)
                pc[k] = __ldcs(reinterpret_cast<const float4 *>(st + ((base ^ piece_addr(k)) + 2 * col16)));
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const int e = ((row_lo + 32 * k) << kB) + 2 * col16;
                *reinterpret_cast<float2 *>(tre + e) = make_float2(pc[k].x, pc[k].z);
                *reinterpret_cast<float2 *>(tim + e) = make_float2(pc[k].y, pc[k].w);
            }
        }
        __syncthreads();

        // ---- gates, up to four at a time ----------------------------------------------------------------------
        for (int g = 0; g < ps.n_groups; ++g) {
            const HbmGroup &gr = ps.group[g];
            switch (gr.n_gates) {
            case 4: run_group<4>(ps, gr, tre, tim, s_coef, base); break;
            case 3: run_group<3>(ps, gr, tre, tim, s_coef, base); break;
            case 2: run_group<2>(ps, gr, tre, tim, s_coef, base); break;
            default: run_group<1>(ps, gr, tre, tim, s_coef, base); break;
            }
            __syncthreads();
        }

        // ---- write back / read out ---------------------------------------------------------------------------
        if (!final_pass) {
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const uint32_t addr = (base ^ piece_addr(k)) + 2 * col16;
                const int e = ((row_lo + 32 * k) << kB) + 2 * col16;
                const float2 r2 = *reinterpret_cast<const float2 *>(tre + e), i2 = *reinterpret_cast<const float2 *>(tim + e);
                float4 v = make_float4(r2.x, i2.x, r2.y, i2.y);
                if (cz) {   // A = identity whenever the CZ ring is used
                    const uint32_t mask = (1u << n) - 1u;
                    const uint32_t k0 = addr, k1 = addr + 1;
                    if (__popc(k0 & (((k0 << 1) | (k0 >> (n - 1))) & mask)) & 1) { v.x = -v.x; v.y = -v.y; }
                    if (__popc(k1 & (((k1 << 1) | (k1 >> (n - 1))) & mask)) & 1) { v.z = -v.z; v.w = -v.w; }
                }
                *reinterpret_cast<float4 *>(st + addr) = v;
            }
        } else {
            // <Z_i> partial sums: sign = parity(z_mask & address) is linear in the piece index, so the thread's
            // 32-term sum is a butterfly of FMAs with +-1 multipliers (one small loop body for every output).
            float pr[kPiecesPerThread][2];
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const int e = ((row_lo + 32 * k) << kB) + 2 * col16;
                const float2 r2 = *reinterpret_cast<const float2 *>(tre + e), i2 = *reinterpret_cast<const float2 *>(tim + e);
                pr[k][0] = fmaf(r2.x, r2.x, i2.x * i2.x);
                pr[k][1] = fmaf(r2.y, r2.y, i2.y * i2.y);
            }
            const uint32_t addr0 = (base ^ ra_lo) + 2 * col16;
#pragma unroll 1
            for (int i = 0; i < a.n_out; ++i) {
                const uint32_t zm = ps.z_mask[i];
                const float s_low = (zm & 1u) ? -1.f : 1.f;
                const float s5 = (__popc(zm & rv5) & 1) ? -1.f : 1.f, s6 = (__popc(zm & rv6) & 1) ? -1.f : 1.f;
                const float s7 = (__popc(zm & rv7) & 1) ? -1.f : 1.f, s8 = (__popc(zm & rv8) & 1) ? -1.f : 1.f;
                float u[kPiecesPerThread];
#pragma unroll
                for (int k = 0; k < 16; ++k) u[k] = fmaf(s_low, pr[k][1], pr[k][0]);
#pragma unroll
                for (int k = 0; k < 8; ++k) u[k] = fmaf(s5, u[2 * k + 1], u[2 * k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) u[k] = fmaf(s6, u[2 * k + 1], u[2 * k]);
#pragma unroll
                for (int k = 0; k < 2; ++k) u[k] = fmaf(s7, u[2 * k + 1], u[2 * k]);
                float sum = fmaf(s8, u[1], u[0]);
                if (__popc(zm & addr0) & 1) sum = -sum;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if (lane == 0) s_z[i][warp] = sum;
            }
            __syncthreads();
            if (threadIdx.x < a.n_out) {
                float s = 0.f;
                for (int ww = 0; ww < kWarps; ++ww) s += s_z[threadIdx.x][ww];
                a.zpart[((size_t)circ * tiles_per_circ + tile_id) * a.n_out + threadIdx.x] = s;
            }
        }
        __syncthreads();
    }
}

// per-circuit gate coefficients (Rot fused with the re-uploaded RY) and layer-0 product vectors
__global__ void hbm_prep_kernel(oqrl_spec sp, const float *__restrict__ params, const float *__restrict__ x, int n_rows,
                                int circ0, int n_circ, float4 *__restrict__ coef, float4 *__restrict__ vec0)
{
    const int n = sp.n_qubits, L = sp.n_layers, Lc = max(L, 1);
    const int per_circ = Lc * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n_circ * per_circ;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / per_circ), rem = (int)(i % per_circ);
        const int l = rem / n, w = rem % n;
        const int gc = circ0 + c;                 // global circuit index = cand * n_rows + row
        const int cand = gc / n_rows, row = gc % n_rows;
        const int f = embed_feature(w, sp.n_feat, sp.feature_map);
        float cs = 1.f, sn = 0.f;
        if (f >= 0) sincosf(0.5f * x[(size_t)row * sp.n_feat + f], &sn, &cs);
        float ar = 1.f, ai = 0.f, br = 0.f, bi = 0.f;
        if (L > 0) {
            const float *p = params + ((size_t)cand * L * n + (size_t)l * n + w) * 3;
            rot_su2(p[0], p[1], p[2], ar, ai, br, bi);
        }
        if (l == 0) {
            // v = [[a, b], [-conj b, conj a]] (cs, sn)^T
            vec0[((size_t)c * n + w) * 2 + 0] = make_float4(ar * cs + br * sn, ai * cs + bi * sn, 0.f, 0.f);
            vec0[((size_t)c * n + w) * 2 + 1] = make_float4(-br * cs + ar * sn, bi * cs - ai * sn, 0.f, 0.f);
        } else if (sp.reupload && f >= 0) {
            // Rot * RY(x): a' = a c + b s, b' = b c - a s
            const float ar2 = ar * cs + br * sn, ai2 = ai * cs + bi * sn;
            const float br2 = br * cs - ar * sn, bi2 = bi * cs - ai * sn;
            ar = ar2; ai = ai2; br = br2; bi = bi2;
        }
        coef[(size_t)c * per_circ + rem] = make_float4(ar, ai, br, bi);
    }
}

__global__ void hbm_zreduce_kernel(const float *__restrict__ zpart, int n_circ, int n_tiles, int n_out,
                                   float *__restrict__ q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_circ * n_out) return;
    const int c = i / n_out, o = i % n_out;
    double s = 0.0;
    for (int t = 0; t < n_tiles; ++t) s += (double)zpart[((size_t)c * n_tiles + t) * n_out + o];
    q[i] = (float)s;
}

// test hook: logical-order copy of the tracked-layout state, out[k] = state[A k]
__global__ void hbm_unpermute_kernel(const float2 *__restrict__ state, const HbmPass *plan, int last_pass, int n,
                                     float2 *__restrict__ out)
{
    __shared__ uint32_t col[32];
    if (threadIdx.x < 32) col[threadIdx.x] = plan[last_pass].a_col[threadIdx.x];
    __syncthreads();
    const size_t dim = (size_t)1 << n;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < dim; k += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = 0;
        for (int p = 0; p < n; ++p)
            if ((k >> p) & 1) x ^= col[p];
        out[k] = state[x];
    }
}

}  // namespace

// workspace layout for a batch of circuits
struct HbmWorkspace {
    size_t plan_off, coef_off, vec0_off, zpart_off, state_off, total;
};
static HbmWorkspace hbm_layout(const oqrl_spec &sp, int n_pass, int batch, bool keep_final_state)
{
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    HbmWorkspace w;
    const int n = sp.n_qubits, Lc = sp.n_layers > 0 ? sp.n_layers : 1;
    size_t off = 0;
    w.plan_off = off; off += al((size_t)n_pass * sizeof(HbmPass));
    w.coef_off = off; off += al((size_t)batch * Lc * n * sizeof(float4));
    w.vec0_off = off; off += al((size_t)batch * n * 2 * sizeof(float4));
    w.zpart_off = off; off += al((size_t)batch * ((size_t)1 << (n - kT)) * sp.n_out * sizeof(float));
    w.state_off = off; off += al(((size_t)batch << n) * sizeof(float2));
    if (keep_final_state) off += al(((size_t)1 << n) * sizeof(float2));
    w.total = off;
    return w;
}

// circuits per batch: bounded so that the states of one batch stay within ~1 GiB
static int hbm_batch(const oqrl_spec &sp, int n_circuits)
{
    const size_t per = ((size_t)1 << sp.n_qubits) * sizeof(float2);
    size_t b = ((size_t)1 << 30) / per;
    if (b < 1) b = 1;
    if (b > (size_t)n_circuits) b = (size_t)n_circuits;
    return (int)b;
}

size_t hbm_workspace_bytes(const oqrl_spec &sp, int n_circuits)
{
    std::vector<HbmPass> plan;
    if (build_plan(sp, plan)) return 0;
    return hbm_layout(sp, (int)plan.size(), hbm_batch(sp, n_circuits), true).total;
}

int hbm_qvalues(const oqrl_spec &sp, const float *d_params, int n_cand, const float *d_x, int n_rows, float *d_q,
                float *d_state_out, void *workspace, size_t workspace_bytes, int sm_count, cudaStream_t st)
{
    std::vector<HbmPass> plan;
    int rc = build_plan(sp, plan, d_state_out != nullptr);
    if (rc) return rc;
    if (d_state_out && n_cand * n_rows != 1) { set_error("statevector hook takes exactly one circuit"); return OQRL_ERR_INVALID; }
    const int n = sp.n_qubits, n_pass = (int)plan.size();
    const int n_circ = n_cand * n_rows;
    const int batch = hbm_batch(sp, n_circ);
    const HbmWorkspace w = hbm_layout(sp, n_pass, batch, d_state_out != nullptr);
    if (w.total > workspace_bytes) { set_error("HBM workspace too small: need %zu, have %zu", w.total, workspace_bytes); return OQRL_ERR_NOMEM; }
    char *ws = (char *)workspace;
    HbmPass *d_plan = (HbmPass *)(ws + w.plan_off);
    OQRL_CUDA_CHECK(cudaMemcpyAsync(d_plan, plan.data(), (size_t)n_pass * sizeof(HbmPass), cudaMemcpyHostToDevice, st));
    const int tiles_per_circ = 1 << (n - kT);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(hbm_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float2) << kT)));
    for (int c0 = 0; c0 < n_circ; c0 += batch) {
        const int nb = n_circ - c0 < batch ? n_circ - c0 : batch;
        HbmArgs a;
        a.plan = d_plan;
        a.state = (float2 *)(ws + w.state_off);
        a.coef = (const float4 *)(ws + w.coef_off);
        a.vec0 = (const float4 *)(ws + w.vec0_off);
        a.zpart = (float *)(ws + w.zpart_off);
        a.n = n; a.n_layers = sp.n_layers; a.n_out = sp.n_out; a.n_circ = nb;
        {
            const long long items = (long long)nb * (sp.n_layers > 0 ? sp.n_layers : 1) * n;
            int blocks = (int)((items + 255) / 256);
            if (blocks > sm_count * 8) blocks = sm_count * 8;
            hbm_prep_kernel<<<blocks, 256, 0, st>>>(sp, d_params, d_x, n_rows, c0, nb, (float4 *)(ws + w.coef_off),
                                                    (float4 *)(ws + w.vec0_off));
            OQRL_LAUNCH_CHECK("hbm_prep_kernel");
        }
        const long long total_tiles = (long long)tiles_per_circ * nb;
        static const int grid_mult = []() { const char *v = getenv("OQRL_HBM_GRID_MULT"); return v ? atoi(v) : 4; }();
        int grid = sm_count * grid_mult;   // persistent CTAs: a multiple of the two resident per SM (3 per SM measured 5 % slower)
        if ((long long)grid > total_tiles) grid = (int)total_tiles;
        for (int p = 0; p < n_pass; ++p) {
            a.pass = p;
            hbm_pass_kernel<<<grid, kHbmThreads, sizeof(float2) << kT, st>>>(a);
            OQRL_LAUNCH_CHECK("hbm_pass_kernel");
        }
        if (d_q) {
            const int items = nb * sp.n_out;
            hbm_zreduce_kernel<<<(items + 127) / 128, 128, 0, st>>>(a.zpart, nb, tiles_per_circ, sp.n_out,
                                                                 d_q + (size_t)c0 * sp.n_out);
            OQRL_LAUNCH_CHECK("hbm_zreduce_kernel");
        }
    }
    if (d_state_out) {
        // test hook (single circuit): the plan was built without the FINAL flag, so the last pass stored the state
        hbm_unpermute_kernel<<<sm_count * 4, 256, 0, st>>>((const float2 *)(ws + w.state_off), d_plan, n_pass - 1, n,
                                                          (float2 *)d_state_out);
        OQRL_LAUNCH_CHECK("hbm_unpermute_kernel");
    }
    return OQRL_OK;
}

double hbm_flops_per_eval(const oqrl_spec &sp)
{
    const int n = sp.n_qubits, L = sp.n_layers;
    const double dim = (double)((size_t)1 << n);
    double f = 6.0 * dim * (double)(n - kB + 1);             // generation: one complex multiply per wire above the row, + 1
    if (L > 1) {
        const unsigned lm = (sp.reserved & 1) ? 0xffffffffu : readout_wire_mask(n, L, sp.n_out, sp.entangler);
        int last_gates = 0;
        for (int w = 0; w < n; ++w) last_gates += (lm >> w) & 1u;
        f += 14.0 * dim * (n * (L - 2) + last_gates);        // SU(2) gates of layers 1..L-1 (re-uploaded RY fused), last layer pruned
    }
    f += 3.0 * dim + (double)sp.n_out * dim;                 // |amp|^2 and signed sums
    return f;
}

double hbm_sweeps_per_eval(const oqrl_spec &sp)
{
    std::vector<HbmPass> plan;
    if (build_plan(sp, plan)) return 0.0;
    return (double)plan.size() - 1.0;                        // each sweep = one read + one write of the 2^n x 8 B state
}

// Test export: the binary pass list (array of HbmPass) of a spec.
int hbm_plan_dump(const oqrl_spec &sp, void *buf, size_t cap, size_t *n_bytes, int keep_state)
{
    std::vector<HbmPass> plan;
    int rc = build_plan(sp, plan, keep_state != 0);
    if (rc) return rc;
    const size_t need = plan.size() * sizeof(HbmPass);
    if (n_bytes) *n_bytes = need;
    if (buf) {
        if (cap < need) { set_error("plan buffer too small"); return OQRL_ERR_INVALID; }
        memcpy(buf, plan.data(), need);
    }
    return OQRL_OK;
}

}  // namespace oqrl
