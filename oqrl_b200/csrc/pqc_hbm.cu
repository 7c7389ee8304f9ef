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
#include <string.h>

#include <array>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace oqrl {

namespace {

constexpr int kT = 13;          // log2 amplitudes per tile (two 32 KiB float planes in shared memory)
constexpr int kB = 4;           // log2 amplitudes per contiguous row (128 bytes)
constexpr int kHbmThreads = 256;
constexpr int kGroups = 3;      // register-block rounds per pass
constexpr int kSlots = 4 * kGroups;   // gate slots per pass; slot s owns shared-memory coordinate bit 1 + s
constexpr int kHbmMaxOut = 8;
static_assert(1 + kSlots == kT, "coordinate bit 0 pairs two blocks, bits 1..12 are the gate slots");

// Shared-memory coordinates.  Inside a tile an amplitude is addressed by kT coordinate bits chosen by the
// planner so that EVERY gate of the pass flips exactly one coordinate bit (slot s <-> bit 1+s) and bit 0 is
// a direction no gate touches.  A register block (4 slots) is then 16 contiguous-stride floats per plane,
// the two blocks a thread packs into fp32x2 differ in bit 0 (adjacent floats -> one 64-bit access), and the
// gate phase needs no address arithmetic beyond one XOR for the lane rotation of group 0.
struct HbmGate {
    int32_t layer, wire;            // wire < 0: empty slot
    uint32_t rho;                   // role functional in shared-memory coordinates (kT bits)
    uint32_t r_phys;                // role functional on the tile base address (n bits)
};
struct HbmPass {
    int32_t flags;                  // PASS_* bits
    int32_t n_tile_bits;            // n - kT
    uint8_t tile_pos[32];           // physical bit positions enumerating the tiles
    uint32_t row_vec[kT - kB];      // address contribution of each row-index bit (low kB bits are zero)
    uint8_t row_piv[kT - kB];       // pivot bit positions of row_vec, ascending: tile base = tile id with zeros inserted there
    uint8_t pad[16 - (kT - kB)];
    int32_t n_gates;                // occupied slots
    int32_t n_groups;               // bit mask: group g (slots 4g..4g+3) holds at least one gate
    HbmGate gate[kSlots];
    uint16_t t_col[kT];             // shared-memory coordinate of tile-relative index bit j (index = row << kB | low)
    uint16_t pad2[3];
    uint32_t z_mask[kHbmMaxOut];    // final pass: <Z_i> sign = parity(z_mask[i] & x)
    uint32_t a_col[32];             // layout after this pass's layer (for the un-permute test hook)
};
enum { PASS_FIRST = 1,     // generate the product state (layer 0 folded in) instead of loading
       PASS_FINAL = 2,     // reduce |amp|^2 into <Z_i> instead of storing
       PASS_CZ = 4,        // CZ-ring sign on the store (this layer's entangler)
       PASS_CZ_PRE = 8 };  // CZ-ring sign right after generation (layer 0's entangler)

inline int parity32(uint32_t v) { return __builtin_popcount(v) & 1; }

// reduced echelon basis over GF(2) with "highest set bit" pivots; every vector carries a tag = which of the
// originally inserted vectors it is the XOR of, so any vector of the span can be expressed in the inserted basis
struct Echelon {
    std::vector<uint32_t> vec, tag;
    std::vector<int> pivot;
    uint32_t reduce(uint32_t v, uint32_t *t = nullptr) const
    {
        for (size_t i = 0; i < vec.size(); ++i)
            if ((v >> pivot[i]) & 1u) { v ^= vec[i]; if (t) *t ^= tag[i]; }
        return v;
    }
    bool add(uint32_t v, uint32_t t = 0)
    {
        v = reduce(v, &t);
        if (!v) return false;
        const int p = 31 - __builtin_clz(v);
        for (size_t i = 0; i < vec.size(); ++i)
            if ((vec[i] >> p) & 1u) { vec[i] ^= v; tag[i] ^= t; }
        vec.push_back(v);
        tag.push_back(t);
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
    const std::array<int, 5> key = {sp.n_qubits, sp.n_layers, sp.n_out, sp.entangler, keep_state ? 1 : 0};
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
    int plan_error = 0;

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
        {
            int np = 0;
            for (int pos = kB; pos < n; ++pos)
                if ((piv >> pos) & 1u) ps.row_piv[np++] = (uint8_t)pos;
        }
        // Shared-memory coordinate basis: bit 1+s = direction of the gate in slot s; bit 0 and the empty slots are
        // filled with role-neutral tile directions.  Coordinate bits 0..4 are the 32 banks a warp touches when the
        // tile is staged (lane bits = physical bits 1..3 of the 16-byte piece + row bits 0, 1), so group 0 (bits
        // 1..4) is given gates whose role functionals are independent on those five lane directions, free group-0
        // slots get fillers built from the lane directions themselves, and a few choices of the two lane row
        // vectors are tried; the assignment with the highest GF(2) rank (5 = conflict free) wins.
        ps.n_gates = (int)gate_pos.size();
        int slot_gate[kSlots];                     // logical bit position of the gate in each slot, -1 = empty
        uint32_t basis[kT];
        auto build_basis = [&](const int *slots, const uint32_t *rv, uint16_t *tcol, uint32_t *bas) -> bool {
            bool have[kT] = {false};
            Echelon co;
            for (int sl = 0; sl < kSlots; ++sl) {
                if (slots[sl] < 0) continue;
                bas[1 + sl] = a_col[slots[sl]];
                have[1 + sl] = true;
                if (!co.add(bas[1 + sl], 1u << (1 + sl))) return false;   // gate directions are independent by construction
            }
            // filler candidates, lane directions first so that they land on the low coordinate bits
            std::vector<uint32_t> cand = {2u, 4u, 8u, rv[0], rv[1], 1u};
            for (int j = 2; j < kT - kB; ++j) cand.push_back(rv[j]);
            // role-neutral: f' = f ^ sum_p (r_p . f) d_p has r_q . f' = 0 for every gate q of the pass, so a gate's
            // role depends on its OWN coordinate bit only (rho = unit vector) and blocks A/B share roles
            for (auto &f : cand)
                for (int sl = 0; sl < kSlots; ++sl)
                    if (slots[sl] >= 0 && parity32(ainv_row[slots[sl]] & f)) f ^= a_col[slots[sl]];
            size_t ci = 0;
            for (int bit = 0; bit < kT; ++bit) {
                if (have[bit]) continue;
                while (ci < cand.size() && !co.add(cand[ci], 1u << bit)) ++ci;
                if (ci >= cand.size()) return false;
                bas[bit] = cand[ci++];
                have[bit] = true;
            }
            for (int j = 0; j < kT; ++j) {
                const uint32_t v = j < kB ? (1u << j) : rv[j - kB];
                uint32_t t = 0;
                if (co.reduce(v, &t) != 0) return false;
                tcol[j] = (uint16_t)t;
            }
            return true;
        };
        auto bank_rank = [&](const uint16_t *tcol) {
            Echelon e;
            int rank = 0;
            const int lane_bits[5] = {1, 2, 3, kB, kB + 1};
            for (int j = 0; j < 5; ++j) rank += e.add(tcol[lane_bits[j]] & 31u) ? 1 : 0;
            return rank;
        };
        {
            uint32_t rowv[kT - kB];
            for (int j = 0; j < kT - kB; ++j) rowv[j] = ps.row_vec[j];
            int best = -1, best_slots[kSlots];
            uint32_t best_rows[kT - kB];
            const int ng = ps.n_gates;
            for (int ra = 0; ra < kT - kB && best < 5; ++ra) {
                for (int rb = 0; rb < kT - kB && best < 5; ++rb) {
                    if (rb == ra) continue;
                    uint32_t rv[kT - kB];
                    int w = 2;
                    rv[0] = rowv[ra]; rv[1] = rowv[rb];
                    for (int j = 0; j < kT - kB; ++j)
                        if (j != ra && j != rb) rv[w++] = rowv[j];
                    int slots[kSlots];
                    for (int sl = 0; sl < kSlots; ++sl) slots[sl] = -1;
                    const int rounds = (ng + 3) / 4;
                    if (rounds <= 2) {
                        // groups 1 and 2 take the gates; group 0 (the bank bits) is left to fillers made from the
                        // lane directions themselves, which makes the staging conflict free by construction
                        for (int i = 0; i < ng; ++i) slots[4 + i] = gate_pos[i];
                    } else {
                        // three rounds: group 0 gets gates whose role restrictions on the lane directions are independent
                        const uint32_t lane_dir[5] = {2u, 4u, 8u, rv[0], rv[1]};
                        std::vector<int> g0, others;
                        Echelon e;
                        for (int i = 0; i < ng; ++i) {
                            uint32_t restr = 0;
                            for (int j = 0; j < 5; ++j) restr |= (uint32_t)parity32(ainv_row[gate_pos[i]] & lane_dir[j]) << j;
                            if (g0.size() < 4 && e.add(restr)) g0.push_back(gate_pos[i]);
                            else others.push_back(gate_pos[i]);
                        }
                        while (others.size() > 8 && g0.size() < 4) { g0.push_back(others.back()); others.pop_back(); }
                        for (size_t i = 0; i < g0.size(); ++i) slots[i] = g0[i];
                        for (size_t i = 0; i < others.size(); ++i) slots[4 + i] = others[i];
                    }
                    uint16_t tcol[kT];
                    uint32_t bas[kT];
                    if (!build_basis(slots, rv, tcol, bas)) continue;
                    const int rk = bank_rank(tcol);
                    if (rk > best) { best = rk; memcpy(best_slots, slots, sizeof(slots)); memcpy(best_rows, rv, sizeof(rv)); }
                }
            }
            if (best < 0) plan_error = 1;
            else {
                memcpy(slot_gate, best_slots, sizeof(best_slots));
                for (int j = 0; j < kT - kB; ++j) ps.row_vec[j] = best_rows[j];
                if (!build_basis(slot_gate, best_rows, ps.t_col, basis)) plan_error = 1;
            }
        }
        ps.n_groups = 0;   // bit mask of the groups that hold at least one gate
        for (int sl = 0; sl < kSlots; ++sl)
            if (!plan_error && slot_gate[sl] >= 0) ps.n_groups |= 1 << (sl / 4);
        for (int s = 0; s < kSlots; ++s) {
            HbmGate &hg = ps.gate[s];
            if (plan_error || slot_gate[s] < 0) { hg.layer = 0; hg.wire = -1; hg.rho = 0; hg.r_phys = 0; continue; }
            const int p = slot_gate[s];
            hg.layer = layer;
            hg.wire = n - 1 - p;
            hg.r_phys = ainv_row[p];
            uint32_t rho = 0;
            for (int j = 0; j < kT; ++j)
                if (parity32(ainv_row[p] & basis[j])) rho |= 1u << j;
            hg.rho = rho;
            if (rho != (1u << (1 + s))) plan_error = 1;   // fillers are role-neutral by construction
        }
        plan.push_back(ps);
    };

    bool cz_pre = false;
    for (int l = 0; l < L; ++l) {
        // layer 0's Rot gates are folded into the generated product state: no pass of its own
        if (l > 0) {
            std::vector<int> remaining;
            for (int p = n - 1; p >= 0; --p) remaining.push_back(p);
            while (!remaining.empty()) {
                Echelon hi;
                std::vector<int> take, rest;
                for (int p : remaining) {
                    const uint32_t h = a_col[p] & ~low_mask;
                    if ((int)take.size() < kSlots && (hi.reduce(h) == 0 || (int)hi.vec.size() < kT - kB)) {
                        hi.add(h);
                        take.push_back(p);
                    } else {
                        rest.push_back(p);
                    }
                }
                // whole register blocks of four gates are the efficient ones: trim the pass to a multiple of
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
    if (plan_error) { set_error("internal: HBM planner could not build a tile basis"); return OQRL_ERR_INVALID; }
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

// One register-block round: group GI owns coordinate bits [1+4 GI, 5+4 GI).  Thread t holds the 16 elements
// that differ in those bits for BOTH values of coordinate bit 0 (blocks A and B, adjacent floats), i.e. 32
// amplitudes as 16 packed re and 16 packed im values: 2 LDS.64 + 2 STS.64 per element pair, immediate offsets.
// Group 0's blocks are 128 bytes apart between lanes; each lane therefore walks its block starting from element
// (lane & 15) -- a rotation that only changes which element plays which role (folded into the swap flags).
// Empty slots of a partially filled group carry the identity (a = 1, b = 0), so the code is branch-free.
template <int GI>
__device__ __forceinline__ void run_group(const HbmPass &ps, float *tre, const float4 *s_coef, uint32_t base)
{
    constexpr int sh = 1 + 4 * GI;
    const uint32_t t = threadIdx.x;
    const uint32_t ca = ((t & ((1u << (sh - 1)) - 1u)) << 1) | ((t >> (sh - 1)) << (sh + 4));   // block A, element 0
    const uint32_t rot = GI == 0 ? (t & 15u) : 0u;
    char *p0 = reinterpret_cast<char *>(tre) + (((size_t)ca | ((size_t)rot << sh)) << 2);          // element `rot`
    constexpr uint32_t kIm = (1u << kT) * 4;                                                      // im plane offset
    u64 vr[16], vi[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        // element index m ^ rot: XOR on the address for the rotated group, plain immediate otherwise
        const char *q = GI == 0 ? reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(p0) ^ ((uintptr_t)m << (sh + 2)))
                                : p0 + ((size_t)m << (sh + 2));
        vr[m] = *reinterpret_cast<const u64 *>(q);
        vi[m] = *reinterpret_cast<const u64 *>(q + kIm);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const HbmGate &hg = ps.gate[4 * GI + i];
        const float4 c = s_coef[4 * GI + i];
        // The role of a register element for this gate is (its own bit i) ^ s, with s from the tile base and the
        // lane rotation (the planner's fillers are role-neutral, so blocks A and B agree).  s = 1 swaps the roles:
        // X U X = [[conj a, -conj b], [b, a]].
        const bool sw = (((__popc(hg.r_phys & base)) ^ (rot >> i)) & 1u) != 0;
        Su2Scalar g;
        g.ar = c.x; g.ai = sw ? -c.y : c.y; g.br = sw ? -c.z : c.z; g.bi = c.w;
#pragma unroll
        for (int m = 0; m < 16; ++m)
            if (((m >> i) & 1) == 0) su2_pair(g, vr[m], vi[m], vr[m | (1 << i)], vi[m | (1 << i)]);
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        char *q = GI == 0 ? reinterpret_cast<char *>(reinterpret_cast<uintptr_t>(p0) ^ ((uintptr_t)m << (sh + 2)))
                          : p0 + ((size_t)m << (sh + 2));
        *reinterpret_cast<u64 *>(q) = vr[m];
        *reinterpret_cast<u64 *>(q + kIm) = vi[m];
    }
}

__global__ void __launch_bounds__(kHbmThreads, 2)
hbm_pass_kernel(HbmArgs a)
{
    extern __shared__ __align__(128) float tile_planes[];   // re[2^kT] then im[2^kT], shared-memory coordinates
    float *tre = tile_planes, *tim = tile_planes + (1 << kT);
    __shared__ HbmPass ps;
    __shared__ float4 s_coef[kSlots];
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
    // Row address = XOR of row_vec over the row bits; shared-memory coordinate = XOR of t_col over the bits of
    // the tile-relative index (row << kB | 2 col16 [+1]).  The thread-fixed parts are hoisted out of the tile loop.
    const int row_lo = threadIdx.x >> 3, col16 = threadIdx.x & 7;
    uint32_t ra_lo = 0, tc_lo = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j)
        if ((row_lo >> j) & 1) { ra_lo ^= ps.row_vec[j]; tc_lo ^= ps.t_col[kB + j]; }
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if ((col16 >> j) & 1) tc_lo ^= ps.t_col[1 + j];
    const uint32_t rv5 = ps.row_vec[5], rv6 = ps.row_vec[6], rv7 = ps.row_vec[7], rv8 = ps.row_vec[8];
    const uint32_t tc9 = ps.t_col[9], tc10 = ps.t_col[10], tc11 = ps.t_col[11], tc12 = ps.t_col[12], tc0 = ps.t_col[0];
    auto piece_addr = [&](int k) {   // address contribution of piece k (before ^ base)
        uint32_t r = ra_lo;
        if (k & 1) r ^= rv5;
        if (k & 2) r ^= rv6;
        if (k & 4) r ^= rv7;
        if (k & 8) r ^= rv8;
        return r;
    };
    auto piece_coord = [&](int k) {  // shared-memory coordinate of the piece's first amplitude (second: ^ tc0)
        uint32_t c = tc_lo;
        if (k & 1) c ^= tc9;
        if (k & 2) c ^= tc10;
        if (k & 4) c ^= tc11;
        if (k & 8) c ^= tc12;
        return c;
    };
    const int n_gen_chunks = (n - kB + kGenChunkBits - 1) / kGenChunkBits;
    int gen_circ = -1;

    for (long long w = blockIdx.x; w < total_tiles; w += gridDim.x) {
        const int circ = (int)(w / tiles_per_circ);
        const uint32_t tile_id = (uint32_t)(w - (long long)circ * tiles_per_circ);
        const uint32_t base = insert_zeros<kT - kB>(tile_id << kB, ps.row_piv, kT - kB);   // zeros at the row pivots
        float2 *st = a.state + ((size_t)circ << n);

        // gate coefficients of this circuit for this pass
        if (threadIdx.x < kSlots) {
            const HbmGate &hg = ps.gate[threadIdx.x];
            s_coef[threadIdx.x] = hg.wire >= 0 ? a.coef[((size_t)circ * max(a.n_layers, 1) + hg.layer) * n + hg.wire]
                                               : make_float4(1.f, 0.f, 0.f, 0.f);   // empty slot = identity
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

        // ---- stage the tile (interleaved complex64 in HBM -> re / im planes in shared-memory coordinates) -------
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
                const uint32_t c0 = piece_coord(k), c1 = c0 ^ tc0;
                tre[c0] = v0.x; tim[c0] = v0.y; tre[c1] = v1.x; tim[c1] = v1.y;
            }
        } else {
            // all 16 pieces of the thread are in flight at once (64 KiB per CTA), coalesced 128-bit loads
            float4 pc[kPiecesPerThread];
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k)
                pc[k] = __ldcs(reinterpret_cast<const float4 *>(st + ((base ^ piece_addr(k)) + 2 * col16)));
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const uint32_t c0 = piece_coord(k), c1 = c0 ^ tc0;
                tre[c0] = pc[k].x; tim[c0] = pc[k].y; tre[c1] = pc[k].z; tim[c1] = pc[k].w;
            }
        }
        __syncthreads();

        // ---- gates: up to three rounds of four ------------------------------------------------------------------
        if (ps.n_groups & 1) { run_group<0>(ps, tre, s_coef, base); __syncthreads(); }
        if (ps.n_groups & 2) { run_group<1>(ps, tre, s_coef, base); __syncthreads(); }
        if (ps.n_groups & 4) { run_group<2>(ps, tre, s_coef, base); __syncthreads(); }

        // ---- write back / read out ---------------------------------------------------------------------------
        if (!final_pass) {
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const uint32_t addr = (base ^ piece_addr(k)) + 2 * col16;
                const uint32_t c0 = piece_coord(k), c1 = c0 ^ tc0;
                float4 v = make_float4(tre[c0], tim[c0], tre[c1], tim[c1]);
                if (cz) {   // A = identity whenever the CZ ring is used
                    const uint32_t mask = (1u << n) - 1u;
                    const uint32_t k0 = addr, k1 = addr + 1;
                    if (__popc(k0 & (((k0 << 1) | (k0 >> (n - 1))) & mask)) & 1) { v.x = -v.x; v.y = -v.y; }
                    if (__popc(k1 & (((k1 << 1) | (k1 >> (n - 1))) & mask)) & 1) { v.z = -v.z; v.w = -v.w; }
                }
                *reinterpret_cast<float4 *>(st + addr) = v;
            }
        } else {
            float z[kHbmMaxOut];
#pragma unroll
            for (int i = 0; i < kHbmMaxOut; ++i) z[i] = 0.f;
#pragma unroll
            for (int k = 0; k < kPiecesPerThread; ++k) {
                const uint32_t addr = (base ^ piece_addr(k)) + 2 * col16;
                const uint32_t c0 = piece_coord(k), c1 = c0 ^ tc0;
                const float p0 = tre[c0] * tre[c0] + tim[c0] * tim[c0], p1 = tre[c1] * tre[c1] + tim[c1] * tim[c1];
#pragma unroll
                for (int i = 0; i < kHbmMaxOut; ++i) {
                    if (i < a.n_out) {
                        z[i] += (__popc(ps.z_mask[i] & addr) & 1) ? -p0 : p0;
                        z[i] += (__popc(ps.z_mask[i] & (addr + 1)) & 1) ? -p1 : p1;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kHbmMaxOut; ++i) {
                if (i < a.n_out) {
                    float s = z[i];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0) s_z[i][warp] = s;
                }
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
        int grid = sm_count * 3;   // 64 KiB tiles: three CTAs per SM
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
    if (L > 1) f += 14.0 * dim * n * (L - 1);                // SU(2) gates of layers 1..L-1 (re-uploaded RY fused)
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
