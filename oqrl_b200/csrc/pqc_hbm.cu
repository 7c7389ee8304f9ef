// HBM-resident PQC kernels for n >= 14 qubits (BASELINE configs 4 and 5).
//
// Reference code replaced: the QNode simulation behind /root/reference/src/agent.py:36-46 for circuits whose
// state (complex64, 2^n amplitudes) no longer fits on chip.  Work is organised as PASSES over the state.  A pass
// cuts the state into TILES of 2^kT amplitudes (64 KiB); a tile is staged in shared memory by the TMA engine
// (cp.async.bulk, one 256-byte row per copy, completion counted on an mbarrier), every pending gate whose pair
// direction lies inside the tile is applied (up to kT - kB + low-bit gates, in register blocks of four), and the
// tile goes back to HBM with bulk stores -- so a layer costs ceil(n / (kT - kB))-ish read+write sweeps instead of n.
//
// One persistent CTA per SM runs a three-stage ring of tile buffers: while the eight consumer warps apply the gates
// of tile i, the four producer warps have the loads of tile i+1 and the stores of tile i-1 in flight.  The only CTA-wide
// synchronisation is the mbarrier hand-off of a buffer and one named barrier between register-block groups.
//
// Arithmetic: an amplitude is ONE packed fp32x2 value (re, im) -- the natural complex64 layout, loaded with a
// single LDS.64.  Blackwell's FFMA2 takes a half-swap (.LO_HI) and a per-half negation on its operands for free, so
// a complex multiply-accumulate is two FFMA2 and an SU(2) gate on an amplitude pair is 2 FMUL2 + 6 FFMA2 with no
// data movement at all (ptxas folds the swaps of cswap() into the operand modifiers; see profiles/r2_hbm_sass.txt).
//
// Entanglers cost NOTHING here.  A CNOT ring is a GF(2)-linear relabelling of basis states, so instead of moving
// amplitudes the planner tracks the map x = A k between the logical index k and the physical address x:
//     gate on logical bit p  ->  pairs (x, x ^ d_p), d_p = A e_p (column p of A);
//     role of x in its pair  ->  logical bit p of k = parity(r_p & x), r_p = row p of A^-1.
// A tile is the affine subspace  base ^ span{e_0..e_{kB-1}, d_p : p in the pass}: it always contains the kB low
// address bits (whole 256-byte rows) and is closed under every pair direction of the pass.  The host planner
// (build_plan) does the linear algebra once per circuit structure; the kernel only XORs precomputed masks.  The CZ
// ring is a sign folded into the last store of a layer.
//
// The first pass generates the embedded product state (with layer 0's Rot gates folded in, as in the other
// families) instead of reading HBM; the last pass reduces |amp|^2 into <Z_i> out of the registers of its last
// register block instead of writing.  Traffic per circuit = (passes - 1) * 2 * 2^n * 8 bytes (hbm_sweeps_per_eval).
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
constexpr int kB = 5;           // log2 amplitudes per contiguous row (256 bytes = one bulk copy)
constexpr int kRowBits = kT - kB;
constexpr int kRows = 1 << kRowBits;
constexpr int kRowBytes = (1 << kB) * 8;
constexpr int kTileBytes = (1 << kT) * 8;
constexpr int kStages = 3;      // tile buffers per CTA: loading / computing / storing
#ifndef OQRL_HBM_CONSUMER_BITS
#define OQRL_HBM_CONSUMER_BITS 8
#endif
constexpr int kThreadBits = OQRL_HBM_CONSUMER_BITS;   // log2 consumer threads: block-index bits supplied by the thread index
constexpr int kConsumerThreads = 1 << kThreadBits;
constexpr int kConsumerWarps = kConsumerThreads / 32; // 8 = two per SM sub-partition (4 warps streaming two blocks each measured 25 % slower)
constexpr int kBlockBits = 4;         // every register block holds 2^4 amplitudes: groups of fewer than four gates pad their directions
constexpr int kProducerWarps = 4;                     // one per SM sub-partition: each moves a quarter of every tile's rows
constexpr int kHbmThreads = kConsumerThreads + 32 * kProducerWarps;
constexpr int kMaxGroups = (kT + 3) / 4;
constexpr int kHbmMaxOut = 8;

struct HbmGate {
    int32_t layer, wire;
    uint32_t delta;     // pair direction in tile coordinates (kT bits: row index bits above kB column bits)
    uint32_t rho;       // role functional in tile coordinates
    uint32_t r_phys;    // role functional on the tile base address (n bits)
    uint32_t d_phys;    // pair direction as a physical address (n bits)
};
struct HbmGroup {
    int32_t n_gates;
    int32_t gate[4];
    // Register blocks: block beta (kT - n_gates bits; thread index below, iteration above) starts at the tile
    // coordinate XOR_j beta_j tvec[j] -- a basis of the subspace on which every gate of the group sees role 0 --
    // shifted per tile onto the coset whose roles are 0 for that tile's base address.  Element m of the block adds
    // the gate directions selected by the bits of m, so element bit i IS the role for gate i and no coefficient ever
    // changes sign.  Blocks always hold 16 amplitudes: delta[] carries the gates' directions first and, for a group
    // of fewer than four gates, role-0 basis vectors as padding (block index = the first kT - 4 entries of tvec).
    // tvec[0..3] (lane bits 0..3) are chosen to differ in tile-coordinate bits 0..3: amplitudes are
    // 8 bytes, a warp's LDS.64 / STS.64 is served per half-warp, and 16 lanes on 16 distinct bank pairs is conflict-free.
    // Where the role-0 subspace cannot reach a low bit (a gate ON that bit), rot_lane[i] names a lane bit whose threads
    // start their blocks from gate i's partner element instead (role 1: the thread flips the signs of ai and br once).
    uint32_t tvec[kT];
    uint8_t rot_lane[4];            // per gate: lane bit 0..3, or 0xff
    uint32_t delta[4], r_phys[4];   // copies of the gates' direction / role functional (one shared-memory hop in the kernel)
};
struct HbmPass {
    int32_t flags;                  // PASS_* bits
    int32_t n_tile_bits;            // n - kT
    uint8_t tile_pos[32];           // physical bit positions enumerating the tiles
    uint32_t row_vec[kRowBits];     // address contribution of each row-index bit (low kB bits are zero)
    int32_t n_gates;
    HbmGate gate[kT];
    int32_t n_groups;
    HbmGroup group[kMaxGroups];
    uint32_t z_mask[kHbmMaxOut];    // final pass: <Z_i> sign = parity(z_mask[i] & x), x the physical address
    uint32_t z_tile[kHbmMaxOut];    // the same functional on tile coordinates (its part inside the tile)
    uint32_t split_phi;             // tile-coordinate functional that vanishes on every gate direction of the pass (0 = none):
                                    // its two cosets are processed by the two halves of the consumer warps independently
    uint32_t a_col[32];             // layout after this pass's layer (for the un-permute test hook)
    uint8_t row_piv[kRowBits];      // pivot bit positions of row_vec, ascending: tile base = tile id with zeros inserted there
    uint8_t pad2[16 - kRowBits];
};
enum { PASS_FIRST = 1,     // generate the product state (layer 0 folded in) instead of loading
       PASS_FINAL = 2,     // reduce |amp|^2 into <Z_i> instead of storing
       PASS_CZ = 4,        // CZ-ring sign on the store (this layer's entangler)
       PASS_CZ_PRE = 8,    // CZ-ring sign right after generation (layer 0's entangler)
       PASS_DBG_NOGATES = 256,
       PASS_DBG_TIMERS = 512,
       PASS_DBG_NOCOPY = 1024,
       PASS_DBG_NOSPLIT = 2048,
       PASS_DBG_STAGGER = 4096 }; // measurement only (OQRL_HBM_DEBUG=16): half 1 of the consumer warps starts every group 500 cycles late // measurement only (OQRL_HBM_DEBUG=8): all consumer warps on one barrier, no split tiles  // measurement only (OQRL_HBM_DEBUG=4): producers hand buffers over without issuing any bulk copy   // measurement only (OQRL_HBM_DEBUG=2): phase timers, see PhaseTimer  // measurement only (OQRL_HBM_DEBUG=1): skip the register-block groups, i.e. time the bare copy pipeline

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
        for (int pos = n - 1; pos >= kB && (int)hi.vec.size() < kRowBits; --pos) hi.add(1u << pos);
        const uint32_t piv = hi.pivot_mask();
        for (int j = 0; j < kRowBits; ++j) ps.row_vec[j] = hi.vec[j];
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
            for (int j = 0; j < kRowBits; ++j)
                if ((d >> hi.pivot[j]) & 1u) { alpha |= 1u << j; d ^= hi.vec[j]; }
            // d is now inside the low kB bits
            hg.delta = (alpha << kB) | (d & low_mask);
            uint32_t rho = ainv_row[p] & low_mask;
            for (int j = 0; j < kRowBits; ++j)
                if (parity32(ainv_row[p] & hi.vec[j])) rho |= 1u << (kB + j);
            hg.rho = rho;
            hg.r_phys = ainv_row[p];
            hg.d_phys = a_col[p];
        }
        // A functional orthogonal to every gate direction of the pass splits the tile into two cosets that no gate of the
        // pass connects: consumer warps 0-3 and 4-7 take one each and only meet when the tile is handed back, so their
        // load / compute / store phases drift apart and overlap instead of running in lockstep.  Prefer row bits.
        ps.split_phi = 0;
        {
            Echelon dspan;
            for (int g = 0; g < ps.n_gates; ++g) dspan.add(ps.gate[g].delta);
            // phi . v = 0 for all v in span  <=>  phi in the orthogonal complement; try unit functionals and pairs from the top
            auto orth = [&](uint32_t phi) {
                for (uint32_t v : dspan.vec)
                    if (parity32(phi & v)) return false;
                return true;
            };
            // Among the orthogonal functionals prefer one that ignores tile-coordinate bits 0..3 (they select the 8-byte
            // bank pair: a half restricted in them could not spread a half-warp over all 16), then the lightest.
            int best_cost = 1 << 30;
            for (uint32_t phi = 1; phi < (1u << kT); ++phi) {
                const int cost = ((phi & 15u) ? 1000 : 0) + __builtin_popcount(phi);
                if (cost < best_cost && orth(phi)) { best_cost = cost; ps.split_phi = phi; }
            }
        }
        // register blocks of up to four gates, sizes balanced (9 gates -> 3+3+3, not 4+4+1); a pass without gates
        // still gets one (empty) group: its blocks of one amplitude carry the generated state to the store / read-out
        ps.n_groups = 0;
        const int want_groups = ps.n_gates > 0 ? (ps.n_gates + 3) / 4 : 1;
        for (int g0 = 0, gi = 0; gi < want_groups; ++gi) {
            HbmGroup &gr = ps.group[ps.n_groups++];
            gr.n_gates = ps.n_gates / want_groups + (gi < ps.n_gates % want_groups ? 1 : 0);
            for (int i = 0; i < gr.n_gates; ++i) gr.gate[i] = g0 + i;
            // kernel of the group's role functionals: start from the unit vectors and clear each role with the gate's
            // own direction (rho_i . delta_j = [i == j], so the corrections do not interfere)
            std::vector<uint32_t> ker;
            for (int pos = 0; pos < kT; ++pos) {
                uint32_t v = 1u << pos;
                for (int i = 0; i < gr.n_gates; ++i)
                    if (parity32(ps.gate[g0 + i].rho & v)) v ^= ps.gate[g0 + i].delta;
                ker.push_back(v);
            }
            // echelon form with LOWEST-set-bit pivots: the vectors that pivot on bits 0..3 go to lane bits 0..3
            std::vector<uint32_t> basis;
            for (int bit = 0; bit < kT; ++bit) {
                size_t k = 0;
                while (k < ker.size() && !((ker[k] >> bit) & 1u)) ++k;
                if (k == ker.size()) continue;
                const uint32_t piv_vec = ker[k];
                ker.erase(ker.begin() + k);
                for (auto &w : ker)
                    if ((w >> bit) & 1u) w ^= piv_vec;
                basis.push_back(piv_vec);
            }
            // (the kT unit vectors span the whole space, so the kernel has exactly kT - n_gates dimensions)
            if (ps.split_phi) {
                // exactly one basis vector may leave the coset of split_phi, and it must sit on thread bit 7 (the warp
                // half); it is taken from behind the four lane vectors so that their low-bit pivots stay untouched
                int js = -1;
                for (int j = 4; j < (int)basis.size() && js < 0; ++j)
                    if (parity32(ps.split_phi & basis[j])) js = j;
                for (int j = 0; j < 4 && j < (int)basis.size() && js < 0; ++j)
                    if (parity32(ps.split_phi & basis[j])) js = j;
                for (int j = 0; j < (int)basis.size(); ++j)
                    if (j != js && parity32(ps.split_phi & basis[j])) basis[j] ^= basis[js];
                const uint32_t u = basis[js];
                basis.erase(basis.begin() + js);
                basis.insert(basis.begin() + (kThreadBits - 1), u);
            }
            for (int j = 0; j < kT; ++j) gr.tvec[j] = j < (int)basis.size() ? basis[j] : 0u;
            // a group of fewer than four gates still runs 16-amplitude blocks: its last role-0 basis vectors become
            // the missing block directions (no gate, no role), the first kT - 4 enumerate the blocks
            // bank spread of a half-warp = GF(2) rank of lane bits 0..3 on tile-coordinate bits 0..3; brute force over
            // the (tiny) space of gate -> lane-bit rotations, keep the best, prefer fewer rotations
            memset(gr.rot_lane, 0xff, sizeof(gr.rot_lane));
            {
                const int G = gr.n_gates;
                int best_rank = -1, best_used = 99;
                int assign[4] = {4, 4, 4, 4};
                int total = 1;
                for (int i = 0; i < G; ++i) total *= 5;      // lane bit 0..3 or "none" (4)
                for (int code = 0; code < total; ++code) {
                    int cc = code, used_n = 0;
                    bool ok = true;
                    uint32_t used = 0;
                    for (int i = 0; i < G; ++i) {
                        assign[i] = cc % 5; cc /= 5;
                        if (assign[i] < 4) {
                            if ((used >> assign[i]) & 1u) ok = false;
                            used |= 1u << assign[i];
                            ++used_n;
                        }
                    }
                    if (!ok) continue;
                    Echelon e;
                    int rank = 0;
                    for (int j = 0; j < 4 && j < kT - G; ++j) {
                        uint32_t col = gr.tvec[j];
                        for (int i = 0; i < G; ++i)
                            if (assign[i] == j) col ^= ps.gate[g0 + i].delta;
                        if (e.add(col & 15u)) ++rank;
                    }
                    if (rank > best_rank || (rank == best_rank && used_n < best_used)) {
                        best_rank = rank; best_used = used_n;
                        for (int i = 0; i < 4; ++i) gr.rot_lane[i] = (i < G && assign[i] < 4) ? (uint8_t)assign[i] : 0xff;
                    }
                }
            }
            for (int i = 0; i < 4; ++i) {
                gr.delta[i] = i < gr.n_gates ? ps.gate[g0 + i].delta : gr.tvec[kT - kBlockBits + (i - gr.n_gates)];
                gr.r_phys[i] = i < gr.n_gates ? ps.gate[g0 + i].r_phys : 0u;
            }
            g0 += gr.n_gates;
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
                    if ((int)take.size() < kT && (hi.reduce(h) == 0 || (int)hi.vec.size() < kRowBits)) {
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
                    const size_t per_pass = (size_t)kRowBits;   // what a later pass can always absorb
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
    for (int i = 0; i < sp.n_out; ++i) {
        const uint32_t zm = ainv_row[n - 1 - i];
        fin.z_mask[i] = zm;
        uint32_t zt = zm & low_mask;
        for (int j = 0; j < kRowBits; ++j)
            if (parity32(zm & fin.row_vec[j])) zt |= 1u << (kB + j);
        fin.z_tile[i] = zt;
    }
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
    const float4 *coef;         // [n_circ][max(L,1)][n][2] SU(2) coefficients {ar, br, -, -}, {ai, ai, bi, bi} (re-upload fused)
    const float4 *vec0;         // [n_circ][n][2]: per-wire 2-vector of the generated product state {re, im, -, -}
    float *zpart;               // [n_circ][n_tiles][kConsumerWarps][n_out] partial <Z_i> sums per tile and consumer warp (final pass)
    unsigned *fault;            // set when an mbarrier wait timed out (a bug, never a data condition)
    int n, n_layers, n_out, n_circ;
    // Copies of this pass's tile geometry in the kernel's constant bank: the producer warp computes every row address
    // from them on the uniform datapath (values read from shared memory are not provably warp-uniform, and a bulk copy
    // with per-lane operands compiles into a 32-step elect/broadcast loop of eight instructions per copy).
    uint32_t row_vec[kRowBits];
    uint32_t row_piv_packed[2];  // row_piv[0..7], one byte each
    int32_t flags, n_tile_bits;
    // ... and of its gate list, for the groups that read their coefficients from constant memory (run_group<.., UNI>):
    // per group whether it does, and the index of its first gate; per gate its record index (layer * n + wire)
    uint8_t group_uni[kMaxGroups], group_g0[kMaxGroups];
    uint16_t gate_lw[16];
    // several circuits per launch whose gates depend on the candidate only (no re-uploading): one record set per
    // CANDIDATE of the batch in constant memory; circuit c of the batch belongs to candidate (c + cand_off) / n_rows
    int32_t cand_recs, cand_off;              // cand_recs = 0: a single record set (index 0)
    uint32_t rows_magic;                      // ceil(2^32 / n_rows)
};

// Coefficient records of ONE circuit in constant memory (copied there device-to-device after hbm_prep_kernel when a
// launch holds a single circuit: config 5, oqrl_statevector).  Indexed by kernel-parameter values only, the loads are
// provably uniform and ptxas keeps the coefficients in UNIFORM registers: each FFMA2 of the group body then reads two
// 64-bit vector operands instead of three (the body alone: 55 against 47.6 TFLOP/s, profiles/r2_gate_rate.txt forms
// 6 / 8).  Every other way of telling ptxas that the coefficients are uniform left them in vector registers.
constexpr int kConstCoefRecs = 1536;                    // gates (2 float4 each): 48 KiB
__constant__ float4 c_hbm_coef[2 * kConstCoefRecs];

#ifndef OQRL_HBM_GEN_UNROLL
#define OQRL_HBM_GEN_UNROLL 4
#endif
constexpr int kGenUnroll = OQRL_HBM_GEN_UNROLL;         // rows of the generated state written per unrolled step
constexpr int kGenChunkBits = 8;                       // generated state: product of per-8-bit-chunk tables
constexpr int kGenMaxChunks = 4;                       // covers n - kB <= 32

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Wait for a phase.  A wait that lasts seconds can only be a protocol bug: record it and trap rather than hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned *fault)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (4LL << 30)) { *fault = 1u; __threadfence_system(); __trap(); }
    }
}
// The same wait for a warp that expects to wait long (a producer waiting for the consumers): try_wait with a
// suspend-time hint parks the warp in hardware until the phase completes instead of re-polling every ~40 cycles --
// ten instructions per poll that the consumer warps of the same sub-partition pay for in issue slots.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, unsigned *fault)
{
#ifdef OQRL_HBM_NO_PARK
    mbar_wait(bar, parity, fault);
#else
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
        if (clock64() - t0 > (4LL << 30)) { *fault = 1u; __threadfence_system(); __trap(); }
    }
#endif
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the bulk store that follows the hand-off)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Phase timers (OQRL_HBM_DEBUG=2; measurement only): one consumer thread and one producer lane add the cycles they
// spend in each phase of the pipeline into 64-bit counters behind the fault word.
struct PhaseTimer {
    long long t, acc[8];
    bool on;
    __device__ __forceinline__ void start(bool enable) { on = enable; for (int i = 0; i < 8; ++i) acc[i] = 0; if (on) t = clock64(); }
    __device__ __forceinline__ void lap(int k) { if (on) { const long long n = clock64(); acc[k] += n - t; t = n; } }
    __device__ __forceinline__ void flush(unsigned *fault, int slot0)
    {
        if (!on) return;
        unsigned long long *d = reinterpret_cast<unsigned long long *>(fault + 2) + slot0;
        for (int i = 0; i < 8; ++i) atomicAdd(d + i, (unsigned long long)acc[i]);
    }
};
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }
// barrier of one half of the consumer warps (ids 2 and 3) when the pass splits its tiles, of all of them otherwise
__device__ __forceinline__ void half_sync(bool split, int half)
{
    if (split) asm volatile("bar.sync %0, %1;" ::"r"(2 + half), "n"(kConsumerThreads / 2) : "memory");
    else consumer_sync();
}

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

// (complex arithmetic on packed (re, im) amplitudes: common.cuh)
__device__ __forceinline__ u64 ld_amp(uint32_t addr)
{
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_amp(uint32_t addr, u64 v)
{
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// CZ-ring sign of basis state x (A = identity whenever the CZ ring is used): (-1)^{#cyclically adjacent 11 pairs}
__device__ __forceinline__ u64 cz_sign(uint32_t x, int n)
{
    const uint32_t mask = (1u << n) - 1u;
    const uint32_t rot = ((x << 1) | (x >> (n - 1))) & mask;
    return (__popc(x & rot) & 1) ? 0x8000000080000000ull : 0ull;
}

struct TileCtx {
    int g;               // index of the group being run (uniform)
    int crec;            // first constant-memory record of this tile's circuit (run_group<.., UNI>)
    uint32_t buf;        // shared-window address of the tile buffer
    uint32_t base;       // physical address of the tile's origin
    const float4 *coef;  // this circuit's coefficient records of the pass's gates (shared memory, [n_gates][2])
    float *zpart;        // this tile's partial <Z_i> (final pass)
};

// Everything a register-block group needs besides the tile data, held in registers.  It depends only on the plan, the
// thread and the tile's address, so it is prepared BEFORE the barrier (or the wait for the tile) that precedes the
// group: every warp reaches that barrier at the same moment, and a setup done after it leaves the FMA pipe idle.
struct GroupRegs {
    uint32_t c;          // tile coordinate of the thread's first block (role 0 for every gate on this tile)
    uint32_t d8[4];      // gate directions as byte offsets
    uint32_t rot;        // bit i: this thread starts from gate i's partner element (signs of ai, br flipped)
    GateOps ops[4];
};

// shift of a group's block origins for the tile at `base`: the directions of the gates whose role functional is odd there
__device__ __forceinline__ uint32_t group_adjust(const HbmGroup &gr, uint32_t base)
{
    uint32_t adj = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)                      // unused slots hold zeros
        if (__popc(gr.r_phys[i] & base) & 1) adj ^= gr.delta[i];
    return adj;
}
__device__ __forceinline__ void prep_addresses(const HbmGroup &gr, const uint32_t *cbase, uint32_t adj, GroupRegs &R)
{
    const uint32_t w = cbase[threadIdx.x];           // per-thread origin and rotation bits, computed once per launch
    R.c = (w & 0xffffu) ^ adj;
    R.rot = w >> 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) R.d8[i] = gr.delta[i] << 3;
}
__device__ __forceinline__ void prep_coefficients(const HbmGroup &gr, const float4 *coef, GroupRegs &R)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i < gr.n_gates) {
            R.ops[i] = load_gate_ops(coef + 2 * (gr.gate[0] + i));     // staged in shared memory by producer 0
            if ((R.rot >> i) & 1) {
                // roles swapped: X U X = [[conj a, -conj b], [b, a]]
                R.ops[i].pai ^= 0x8000000080000000ull;
                R.ops[i].br = -R.ops[i].br;
            }
        }
    }
}

// Register blocks of 2^G amplitudes (see HbmGroup): element m of a block sits at tile coordinate c0 ^ off[m], off[m]
// the XOR of the gate directions selected by the bits of m, c0 the block's role-0 element.  `last` = this is the
// pass's last group: it stores with the layer's CZ sign, or -- on the final pass -- reduces <Z_i> out of its registers
// instead of storing.
// MODE (compile time, so that a pass only ever executes the code it needs): 0 = store in place, 1 = store with the
// layer's CZ sign, 2 = final pass, reduce <Z_i> out of the registers instead of storing.  Blocks always hold 16
// amplitudes; NG = the group's number of gates (compile time: a run-time test between the gates splits the body into
// basic blocks and cost 5 % on the four-gate groups), the remaining block directions are padding.
// UNI: the group's coefficients come from constant memory (single circuit per launch, no thread of the group starts
// from a partner element -- every group whose gates act above the lane-distribution bits, i.e. the 8-gate passes).
template <int MODE, int NG, bool UNI = false>
__device__ __forceinline__ void run_group(const HbmPass &ps, const HbmGroup &gr, const HbmArgs &a, const TileCtx &tc,
                                          const GroupRegs &R, float (&zacc)[kHbmMaxOut])
{
    constexpr int G = kBlockBits, kElems = 1 << G;
    GateOps ops[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i >= NG) continue;
        if constexpr (UNI) ops[i] = load_gate_ops(c_hbm_coef + 2 * (tc.crec + a.gate_lw[a.group_g0[tc.g] + i]));
        else ops[i] = R.ops[i];
    }
    constexpr int kIterBits = kT - G - kThreadBits;      // block-index bits above the thread-index bits
    static_assert(kIterBits >= 0, "at least one block per thread");
    uint32_t off[kElems];      // element offsets in bytes (XOR and scaling commute)
    off[0] = 0;
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
        for (int m = 0; m < (1 << i); ++m) off[m | (1 << i)] = off[m] ^ R.d8[i];
    constexpr bool final_pass = MODE == 2, cz = MODE == 1;
#pragma unroll 1
    for (int it = 0; it < (1 << kIterBits); ++it) {
        uint32_t c0 = R.c;
#pragma unroll
        for (int j = 0; j < kIterBits; ++j)
            if ((it >> j) & 1) c0 ^= gr.tvec[kThreadBits + j];
        const uint32_t rel0 = c0 << 3;          // element m sits at byte (rel0 ^ off[m]) of the tile buffer
        u64 v[kElems];
#pragma unroll
        for (int m = 0; m < kElems; ++m) v[m] = ld_amp(tc.buf + (rel0 ^ off[m]));
#pragma unroll
        for (int i = 0; i < NG; ++i) {
#pragma unroll
            for (int m = 0; m < kElems; ++m)
                if (((m >> i) & 1) == 0) su2_cpair(ops[i], v[m], v[m | (1 << i)]);
        }
        if (!final_pass) {
            if (cz) {
                // physical address of element 0: tile origin ^ row address ^ column; elements add the block directions
                uint32_t x0 = tc.base ^ (c0 & ((1u << kB) - 1u));
#pragma unroll
                for (int j = 0; j < kRowBits; ++j)
                    if ((c0 >> (kB + j)) & 1u) x0 ^= a.row_vec[j];
                uint32_t offx[kElems];
                offx[0] = 0;
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const uint32_t d = R.d8[i] >> 3;
                    uint32_t dx = d & ((1u << kB) - 1u);
#pragma unroll
                    for (int j = 0; j < kRowBits; ++j)
                        if ((d >> (kB + j)) & 1u) dx ^= a.row_vec[j];
#pragma unroll
                    for (int m = 0; m < (1 << i); ++m) offx[m | (1 << i)] = offx[m] ^ dx;
                }
#pragma unroll
                for (int m = 0; m < kElems; ++m) v[m] ^= cz_sign(x0 ^ offx[m], a.n);
            }
#pragma unroll
            for (int m = 0; m < kElems; ++m) st_amp(tc.buf + (rel0 ^ off[m]), v[m]);
        } else {
            // <Z_i> partial sums: sign = parity(z_mask & address) is linear in the element index, so the block's
            // 16-term sum is a 4-level butterfly of FMAs with +-1 multipliers (one small loop body per output).
            float p[kElems];
#pragma unroll
            for (int m = 0; m < kElems; ++m) {
                float re, im;
                unpack2(v[m], re, im);
                p[m] = fmaf(re, re, im * im);
            }
#pragma unroll 1
            for (int o = 0; o < a.n_out; ++o) {
                const uint32_t zt = ps.z_tile[o];
                float u[kElems];
#pragma unroll
                for (int m = 0; m < kElems; ++m) u[m] = p[m];
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const float sg = (__popc(zt & (R.d8[i] >> 3)) & 1) ? -1.f : 1.f;
#pragma unroll
                    for (int m = 0; m < (kElems >> (i + 1)); ++m) u[m] = fmaf(sg, u[2 * m + 1], u[2 * m]);
                }
                const bool neg = ((__popc(zt & c0) ^ __popc(ps.z_mask[o] & tc.base)) & 1) != 0;
                zacc[o] += neg ? -u[0] : u[0];
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ void run_group_n(const HbmPass &ps, const HbmGroup &gr, const HbmArgs &a, const TileCtx &tc,
                                            const GroupRegs &R, float (&zacc)[kHbmMaxOut])
{
    const bool uni = a.group_uni[tc.g] != 0;         // coefficients from constant memory (uniform registers)
    switch (gr.n_gates) {
    case 4: if (uni) run_group<MODE, 4, true>(ps, gr, a, tc, R, zacc); else run_group<MODE, 4>(ps, gr, a, tc, R, zacc); break;
    case 3: if (uni) run_group<MODE, 3, true>(ps, gr, a, tc, R, zacc); else run_group<MODE, 3>(ps, gr, a, tc, R, zacc); break;
    case 2: if (uni) run_group<MODE, 2, true>(ps, gr, a, tc, R, zacc); else run_group<MODE, 2>(ps, gr, a, tc, R, zacc); break;
    case 1: if (uni) run_group<MODE, 1, true>(ps, gr, a, tc, R, zacc); else run_group<MODE, 1>(ps, gr, a, tc, R, zacc); break;
    default: run_group<MODE, 0>(ps, gr, a, tc, R, zacc); break;
    }
}

// Shared memory: [kStages][64 KiB] tile buffers at the start of the dynamic allocation, then the bookkeeping block.
struct HbmShared {
    HbmPass ps;
    unsigned long long full[kStages];     // producer -> consumers: the tile has landed (or the buffer is free to generate into)
    unsigned long long done[kStages];     // consumers -> producer: gates applied, buffer ready to store / reuse
    float2 gen[kGenMaxChunks][1 << kGenChunkBits];   // partial products of the generated state, per 8-bit address chunk
    uint32_t cbase[kMaxGroups][kConsumerThreads];   // per thread and group: block origin (low 16 bits) | rotation bits << 16
    float4 coef[2][kT][2];                // gate coefficients of the pass for the current / next circuit (staged by producer 0)
    float2 vec0[64];                      // this circuit's per-wire 2-vectors
    // per ring stage, written by producer 0 before it arrives on `full`: the tile's origin and, per group, the shift of
    // the block origins onto the coset whose roles are 0 for this tile (what every consumer thread used to recompute)
    uint32_t tile_base[kStages];
    uint32_t tile_adj[kStages][kMaxGroups];
    float2 rowf[kRows];                   // generated pass: row factors of the tile being written (each warp reads back its own 32)
};
constexpr size_t kHbmSmemBytes = (size_t)kStages * kTileBytes + sizeof(HbmShared) + 1024;

__global__ void __launch_bounds__(kHbmThreads, 1)
hbm_pass_kernel(HbmArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;                  // tile buffers first, 1 KiB-aligned
    HbmShared &sh = *reinterpret_cast<HbmShared *>(smem_raw + (ring - smem_u32(smem_raw)) + (size_t)kStages * kTileBytes);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (role dispatch)
    const int lane = threadIdx.x & 31;

    {   // plan of this pass -> shared memory
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.plan + a.pass);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.ps);
        for (int i = threadIdx.x; i < (int)(sizeof(HbmPass) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&sh.full[s]), kProducerWarps);
            mbar_init(smem_u32(&sh.done[s]), kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < kConsumerThreads) {
        for (int g = 0; g < sh.ps.n_groups; ++g) {
            // tile-independent part of every group's block origin: this thread's share of the role-0 subspace, and
            // whether its blocks start from a gate's partner element (HbmGroup::rot_lane)
            const HbmGroup &gr = sh.ps.group[g];
            uint32_t c = 0, rot = 0;
#pragma unroll
            for (int j = 0; j < kThreadBits; ++j)
                if ((threadIdx.x >> j) & 1) c ^= gr.tvec[j];
            for (int i = 0; i < gr.n_gates; ++i) {
                const int rl = gr.rot_lane[i];
                if (rl != 0xff && ((threadIdx.x >> rl) & 1)) { c ^= gr.delta[i]; rot |= 1u << i; }
            }
            sh.cbase[g][threadIdx.x] = c | (rot << 16);        // read back by the same thread only
        }
    }
    // Programmatic dependent launch: everything above only reads the plan (uploaded before the first pass); from here on
    // the kernel touches what the previous pass / the preparation kernel wrote.  The next pass may be scheduled as soon
    // as this grid's CTAs retire, its launch latency and prologue hidden under this pass's tail.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const HbmPass &ps = sh.ps;
    const int n = a.n;
    const int tiles_per_circ = 1 << ps.n_tile_bits;
    const long long total_tiles = (long long)tiles_per_circ * a.n_circ;
    // contiguous tile range of this CTA (a circuit's tiles stay together: tables are rebuilt only when it changes)
    const long long w0 = total_tiles * blockIdx.x / gridDim.x, w1 = total_tiles * (blockIdx.x + 1) / gridDim.x;
    const int count = (int)(w1 - w0);
    const bool first = ps.flags & PASS_FIRST, final_pass = ps.flags & PASS_FINAL;
    const bool cz_pre = ps.flags & PASS_CZ_PRE;
    if (warp >= kConsumerWarps) {
        // ================= producer warps: TMA loads one tile ahead, TMA stores one tile behind ==================
        // Every lane runs this code with identical, warp-uniform values (kernel parameters, block index, loop counters),
        // so addresses live in uniform registers; one elected lane issues the copies and owns their bulk groups.
        // Producer q moves rows [64 q, 64 q + 64) of every tile -- it loads exactly the rows it stored, so its own
        // wait_group.read frees exactly what it overwrites and the four warps never synchronise with each other.
        const bool leader = elect_one();
        const int q = warp - kConsumerWarps;
        constexpr int kHiPer = kRows / 16 / kProducerWarps;
        const int p_tiles_per_circ = 1 << a.n_tile_bits;
        const bool p_first = a.flags & PASS_FIRST, p_final = a.flags & PASS_FINAL;
        auto p_tile_base = [&](long long w, int &circ) {
            circ = (int)(w >> a.n_tile_bits);
            uint32_t v = ((uint32_t)w & (uint32_t)(p_tiles_per_circ - 1)) << kB;
#pragma unroll
            for (int i = 0; i < kRowBits; ++i) {
                const int p = (a.row_piv_packed[i >> 2] >> (8 * (i & 3))) & 0xff;
                v = ((v >> p) << (p + 1)) | (v & ((1u << p) - 1u));
            }
            return v;
        };
        // row r = 16 hi + lo sits at base ^ xh(hi) ^ xl[lo]
        uint32_t xl[16];
        xl[0] = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int m = 0; m < (1 << b); ++m) xl[m | (1 << b)] = xl[m] ^ a.row_vec[b];
        // Refill order: as soon as tile t has been handed back (done) its store is issued, and once that store has
        // READ the buffer (wait_group.read -- the write to HBM continues in the background) tile t + kStages is loaded
        // into the same buffer.  So two loads are in flight while a third tile computes: one 64 KiB load in flight per
        // SM is not enough to cover the HBM latency at full bandwidth (first build: 0.53 of peak, copy pipeline alone 0.84).
        int staged_circ = -1, n_changes = 0;
        auto issue_load = [&](int t) {
            const int s = t % kStages;
            const uint32_t full = smem_u32(&sh.full[s]);
            int circ;
            const uint32_t base = p_tile_base(w0 + t, circ);
            if (q == 0) {
                // the tile's mailbox (the leader's arrive on `full` below releases it to the consumers)
                if (lane < ps.n_groups) sh.tile_adj[s][lane] = group_adjust(ps.group[lane], base);
                if (lane == 0) sh.tile_base[s] = base;
                __syncwarp();
            }
            if (q == 0 && circ != staged_circ) {
                // gate coefficients of this pass for the tile's circuit -> shared memory, one circuit ahead of the consumers
                staged_circ = circ;
                float4 *dst = &sh.coef[n_changes & 1][0][0];
                ++n_changes;
                const float4 *src = a.coef + 2 * (size_t)circ * max(a.n_layers, 1) * n;
                for (int e = lane; e < 2 * ps.n_gates; e += 32) {
                    const HbmGate &hg = ps.gate[e >> 1];
                    dst[e] = __ldg(src + 2 * ((size_t)hg.layer * n + hg.wire) + (e & 1));
                }
                __syncwarp();            // the leader's arrive below releases these stores to the consumers
            }
            if (p_first || (a.flags & PASS_DBG_NOCOPY)) {
                if (leader) mbar_arrive(full);                       // nothing to load: the buffer is free to generate into
            } else {
                const unsigned long long st = (unsigned long long)a.state + ((unsigned long long)circ << (n + 3));
                if (leader) mbar_arrive_expect_tx(full, kTileBytes / kProducerWarps);
                const uint32_t dst = ring + (uint32_t)s * kTileBytes;
#pragma unroll 1
                for (int hi = q * kHiPer; hi < (q + 1) * kHiPer; ++hi) {
                    uint32_t bh = base;
#pragma unroll
                    for (int b = 0; b < kRowBits - 4; ++b)
                        if ((hi >> b) & 1) bh ^= a.row_vec[4 + b];
#pragma unroll
                    for (int lo = 0; lo < 16; ++lo)
                        if (leader) bulk_load(dst + (uint32_t)(hi * 16 + lo) * kRowBytes,
                                              (const void *)(st + ((unsigned long long)(bh ^ xl[lo]) << 3)), kRowBytes, full);
                }
            }
        };
        PhaseTimer pt;
        pt.start((a.flags & PASS_DBG_TIMERS) && q == 0 && leader);
        for (int t = 0; t < kStages && t < count; ++t) issue_load(t);
        pt.lap(0);
        for (int it = 0; it < count; ++it) {
            const int s = it % kStages;
            mbar_wait_parked(smem_u32(&sh.done[s]), (it / kStages) & 1, a.fault);
            pt.lap(1);
            if (!p_final && !(a.flags & PASS_DBG_NOCOPY)) {
                int circ;
                const uint32_t base = p_tile_base(w0 + it, circ);
                const unsigned long long st = (unsigned long long)a.state + ((unsigned long long)circ << (n + 3));
                const uint32_t src = ring + (uint32_t)s * kTileBytes;
#pragma unroll 1
                for (int hi = q * kHiPer; hi < (q + 1) * kHiPer; ++hi) {
                    uint32_t bh = base;
#pragma unroll
                    for (int b = 0; b < kRowBits - 4; ++b)
                        if ((hi >> b) & 1) bh ^= a.row_vec[4 + b];
#pragma unroll
                    for (int lo = 0; lo < 16; ++lo)
                        if (leader) bulk_store((void *)(st + ((unsigned long long)(bh ^ xl[lo]) << 3)),
                                               src + (uint32_t)(hi * 16 + lo) * kRowBytes, kRowBytes);
                }
                if (leader) bulk_commit();
                pt.lap(2);
                if (it + kStages < count && leader) bulk_wait_read<0>();    // this producer's rows have left the buffer
                pt.lap(3);
            }
            if (it + kStages < count) issue_load(it + kStages);
            pt.lap(4);
        }
        if (!p_final && leader) bulk_wait_all<0>();     // shared memory must outlive the reads of the last stores
        pt.lap(5);
        pt.flush(a.fault, 8);
        return;
    }

    // ===================== consumer warps: generate / apply gates / read out ========================================
    const int n_gen_chunks = (n - kB + kGenChunkBits - 1) / kGenChunkBits;
    int gen_circ = -1, coef_circ = -1, n_changes = 0;
    u64 low_amp = 0ull;                     // this lane's column factor of the generated state
    const bool no_gates = a.flags & PASS_DBG_NOGATES;
    // The two halves of the consumer warps work on the two cosets of split_phi, which no gate of the pass connects:
    // they synchronise among themselves only.  Half 1 starts a few hundred cycles late, once; from then on its loads
    // and stores fall under half 0's arithmetic and vice versa (they re-align only when both wait for the same tile,
    // i.e. when the pass is memory-bound and there is slack anyway).
    const bool split = ps.split_phi != 0 && !(a.flags & PASS_DBG_NOSPLIT);
    const int half = warp >> 2;
    if (split && half == 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < 600) { }
    }
    PhaseTimer ct;
    ct.start((a.flags & PASS_DBG_TIMERS) && threadIdx.x == 0);
    for (int it = 0; it < count; ++it) {
        const int s = it % kStages;
        TileCtx tc;
        const int circ = (int)((w0 + it) >> a.n_tile_bits);
        tc.buf = ring + (uint32_t)s * kTileBytes;
        if (circ != coef_circ) { coef_circ = circ; ++n_changes; }     // same count as producer 0 keeps
        tc.coef = &sh.coef[(n_changes - 1) & 1][0][0];
        tc.g = 0;
        // candidate of the batch = (circ + cand_off) / n_rows as a multiply-high by ceil(2^32 / n_rows) (exact in the host-checked
        // range; a division instruction sequence is where ptxas stops treating the index as uniform)
        tc.crec = a.cand_recs > 0 ? (int)__umulhi((unsigned)(circ + a.cand_off), a.rows_magic) * a.cand_recs : 0;
        const uint32_t tile_id = (uint32_t)(w0 + it) & (uint32_t)(tiles_per_circ - 1);
        tc.zpart = a.zpart + ((size_t)circ * tiles_per_circ + tile_id) * kConsumerWarps * a.n_out;
        GroupRegs R;

        if (first && circ != gen_circ) {
            // tables of the generated product state for this circuit (A = identity while the first pass runs)
            consumer_sync();                                       // the previous circuit's tables are no longer read
            const float4 *vec0 = a.vec0 + (size_t)circ * n * 2;
            for (int e = threadIdx.x; e < 2 * n; e += kConsumerThreads) { const float4 f = vec0[e]; sh.vec0[e] = make_float2(f.x, f.y); }
            consumer_sync();
            {
                u64 p = pack2(1.f, 0.f);
                for (int pos = 0; pos < kB; ++pos) {
                    const float2 f = sh.vec0[(n - 1 - pos) * 2 + ((lane >> pos) & 1)];
                    p = cmulp(p, pack2(f.x, f.y));
                }
                low_amp = p;
            }
            for (int i = threadIdx.x; i < n_gen_chunks << kGenChunkBits; i += kConsumerThreads) {
                const int ch = i >> kGenChunkBits, v = i & ((1 << kGenChunkBits) - 1);
                u64 p = pack2(1.f, 0.f);
                for (int b = 0; b < kGenChunkBits; ++b) {
                    const int pos = kB + ch * kGenChunkBits + b;
                    if (pos < n) {
                        const float2 f = sh.vec0[(n - 1 - pos) * 2 + ((v >> b) & 1)];
                        p = cmulp(p, pack2(f.x, f.y));
                    }
                }
                float re, im;
                unpack2(p, re, im);
                sh.gen[ch][v] = make_float2(re, im);
            }
            gen_circ = circ;
            consumer_sync();
        }

        ct.lap(0);                                                   // tile setup
        mbar_wait(smem_u32(&sh.full[s]), (it / kStages) & 1, a.fault);
        ct.lap(1);                                                   // waiting for the tile
        tc.base = sh.tile_base[s];                                   // mailbox of producer 0: origin and block-origin shifts
        prep_addresses(ps.group[0], sh.cbase[0], sh.tile_adj[s][0], R);
        if (!a.group_uni[0]) prep_coefficients(ps.group[0], tc.coef, R);   // staged by producer 0 before it arrived on `full`; a
                                                                             // group served from constant memory needs none

        if (first) {
            // warp w writes kRows / kConsumerWarps rows, 32 at a time, lane = column: one coalesced 256-byte row per
            // STS.64.  Lane j prepares the row factor of row r0 + j (chunk-table look-ups) and the warp shuffles it
            // round; computing the row factor per element instead measured 4x slower (first pass of config 4).
#pragma unroll 1
            for (int r0 = warp * (kRows / kConsumerWarps); r0 < (warp + 1) * (kRows / kConsumerWarps); r0 += 32) {
                u64 rowp;
                {
                    const int r = r0 + lane;
                    uint32_t x = tc.base;
#pragma unroll
                    for (int b = 0; b < kRowBits; ++b)
                        if ((r >> b) & 1) x ^= a.row_vec[b];
                    u64 p = pack2(1.f, 0.f);
                    for (int ch = 0; ch < n_gen_chunks; ++ch) {
                        const float2 f = sh.gen[ch][(x >> (kB + ch * kGenChunkBits)) & ((1 << kGenChunkBits) - 1)];
                        p = ch == 0 ? pack2(f.x, f.y) : cmulp(p, pack2(f.x, f.y));
                    }
                    rowp = p;
                }
                // two copies of the loop: a test of cz_pre inside it ends a basic block per row, and the rows' shuffle ->
                // multiply -> store chains then run one after the other (2.5 k cycles per tile instead of ~0.6 k)
                if (!cz_pre) {
                    // The row factors go through shared memory (each lane parks the one it computed, the warp reads them
                    // back as broadcast LDS.64) instead of 64 shuffles per warp: eight warps shuffling at once are bound by
                    // the SM's shuffle rate.  Eight rows at a time: loads and stores are ordered among themselves
                    // (volatile), so interleaving them row by row would leave every row waiting out its own load latency.
                    const uint32_t rowf = smem_u32(&sh.rowf[r0]);
                    {
                        float re, im;
                        unpack2(rowp, re, im);
                        sh.rowf[r0 + lane] = make_float2(re, im);
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int j0 = 0; j0 < 32; j0 += 8) {
                        u64 v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = cmulp(ld_amp(rowf + (uint32_t)(j0 + k) * 8u), low_amp);
#pragma unroll
                        for (int k = 0; k < 8; ++k) st_amp(tc.buf + (uint32_t)(r0 + j0 + k) * kRowBytes + (uint32_t)lane * 8u, v[k]);
                    }
                } else {
#pragma unroll kGenUnroll
                    for (int j = 0; j < 32; ++j) {
                        const u64 rp = __shfl_sync(0xffffffffu, rowp, j);
                        u64 v = cmulp(rp, low_amp);
                        const int r = r0 + j;
                        uint32_t x = tc.base ^ (uint32_t)lane;
#pragma unroll
                        for (int b = 0; b < kRowBits; ++b)
                            if ((r >> b) & 1) x ^= a.row_vec[b];
                        v ^= cz_sign(x, n);
                        st_amp(tc.buf + (uint32_t)r * kRowBytes + (uint32_t)lane * 8u, v);
                    }
                }
            }
            ct.lap(7);                                               // generation: rows written
            consumer_sync();
            ct.lap(2);                                               // generation: waiting for the other warps
        }

        // ---- gates, up to four at a time ----------------------------------------------------------------------
        float zacc[kHbmMaxOut];
#pragma unroll
        for (int o = 0; o < kHbmMaxOut; ++o) zacc[o] = 0.f;
        if (!no_gates) {
#pragma unroll 1
            for (int g = 0; g < ps.n_groups; ++g) {
                const bool last = g == ps.n_groups - 1;
                tc.g = g;
                if ((a.flags & PASS_DBG_STAGGER) && half == 1) {
                    const long long t0 = clock64();
                    while (clock64() - t0 < 500) { }
                }
                const int mode = !last ? 0 : (final_pass ? 2 : ((ps.flags & PASS_CZ) ? 1 : 0));     // one call site per mode: one copy of each body
                if (mode == 0) run_group_n<0>(ps, ps.group[g], a, tc, R, zacc);
                else if (mode == 2) run_group_n<2>(ps, ps.group[g], a, tc, R, zacc);
                else run_group_n<1>(ps, ps.group[g], a, tc, R, zacc);
                ct.lap(3);                                           // register-block group
                if (!last) {
                    // the next group's setup goes in front of the barrier, where it overlaps the other warps' tails
                    prep_addresses(ps.group[g + 1], sh.cbase[g + 1], sh.tile_adj[s][g + 1], R);
                    if (!a.group_uni[g + 1]) prep_coefficients(ps.group[g + 1], tc.coef, R);
                    ct.lap(4);                                       // next group's setup
                    half_sync(split, half);
                    ct.lap(5);                                       // barrier
                }
            }
        }

        if (final_pass) {
            // one partial sum per tile and consumer WARP, written by the warp itself (no barrier, no shared-memory hop: the
            // per-half sums this replaces cost a barrier per tile, 0.76 k cycles of the final pass's 4.2 k), added up in a
            // fixed order by hbm_zreduce_kernel
            for (int o = 0; o < a.n_out; ++o) {
                float v = zacc[o];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) tc.zpart[warp * a.n_out + o] = v;
            }
        } else {
            fence_async_smem();          // this thread's stores -> visible to the bulk store issued by the producer
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sh.done[s]));
        ct.lap(6);                                                   // read-out / fence / hand-off
    }
    ct.flush(a.fault, 0);
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
        coef[2 * ((size_t)c * per_circ + rem)] = make_float4(ar, br, 0.f, 0.f);
        coef[2 * ((size_t)c * per_circ + rem) + 1] = make_float4(ai, ai, bi, bi);
    }
}

// <Z_i> of a circuit = the sum of its per-tile, per-warp partial sums.  One block per (circuit, output): thread t adds
// the entries t, t + 128, ... in that order (fp64), the block adds the 128 thread sums in a fixed tree -- the same order
// whatever the batch, so a value's bits do not depend on the launch.  (One THREAD per value, as this kernel first was,
// walked 4096 entries serially for a 24-qubit circuit: 0.1 ms of a 1.3 ms circuit.)
constexpr int kZReduceThreads = 128;
__global__ void __launch_bounds__(kZReduceThreads)
hbm_zreduce_kernel(const float *__restrict__ zpart, int n_circ, int n_tiles, int n_out, float *__restrict__ q)
{
    __shared__ double part[kZReduceThreads];
    const int c = blockIdx.x / n_out, o = blockIdx.x - c * n_out;
    const int n_part = kConsumerWarps * n_tiles;                 // one per tile and warp
    const float *z = zpart + (size_t)c * n_part * n_out + o;
    double s = 0.0;
    for (int t = threadIdx.x; t < n_part; t += kZReduceThreads) s += (double)z[(size_t)t * n_out];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int w = kZReduceThreads / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) q[blockIdx.x] = (float)part[0];
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
    size_t plan_off, coef_off, vec0_off, zpart_off, fault_off, state_off, total;
};
static HbmWorkspace hbm_layout(const oqrl_spec &sp, int n_pass, int batch, bool keep_final_state)
{
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    HbmWorkspace w;
    const int n = sp.n_qubits, Lc = sp.n_layers > 0 ? sp.n_layers : 1;
    size_t off = 0;
    w.plan_off = off; off += al((size_t)n_pass * sizeof(HbmPass));
    w.coef_off = off; off += al((size_t)batch * Lc * n * 2 * sizeof(float4));
    w.vec0_off = off; off += al((size_t)batch * n * 2 * sizeof(float4));
    w.zpart_off = off; off += al((size_t)batch * ((size_t)kConsumerWarps << (n - kT)) * sp.n_out * sizeof(float));
    w.fault_off = off; off += 256;
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
    OQRL_CUDA_CHECK(cudaMemsetAsync(ws + w.fault_off, 0, 256, st));
    const int tiles_per_circ = 1 << (n - kT);
    OQRL_CUDA_CHECK(cudaFuncSetAttribute(hbm_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHbmSmemBytes));
    for (int c0 = 0; c0 < n_circ; c0 += batch) {
        const int nb = n_circ - c0 < batch ? n_circ - c0 : batch;
        HbmArgs a;
        a.plan = d_plan;
        a.state = (float2 *)(ws + w.state_off);
        a.coef = (const float4 *)(ws + w.coef_off);
        a.vec0 = (const float4 *)(ws + w.vec0_off);
        a.zpart = (float *)(ws + w.zpart_off);
        a.fault = (unsigned *)(ws + w.fault_off);
        a.n = n; a.n_layers = sp.n_layers; a.n_out = sp.n_out; a.n_circ = nb;
        {
            const long long items = (long long)nb * (sp.n_layers > 0 ? sp.n_layers : 1) * n;
            int blocks = (int)((items + 255) / 256);
            if (blocks > sm_count * 8) blocks = sm_count * 8;
            hbm_prep_kernel<<<blocks, 256, 0, st>>>(sp, d_params, d_x, n_rows, c0, nb, (float4 *)(ws + w.coef_off),
                                                    (float4 *)(ws + w.vec0_off));
            OQRL_LAUNCH_CHECK("hbm_prep_kernel");
        }
        const int recs = (sp.n_layers > 0 ? sp.n_layers : 1) * n;
        // Coefficients through constant memory (run_group<.., UNI>): a single circuit, or -- gates that depend on the
        // candidate only -- one record set per candidate of the batch (circuits are candidate-major: c = cand * n_rows + row)
        const int cand_off = c0 % n_rows;
        const int n_cands = (cand_off + nb + n_rows - 1) / n_rows;
        const bool no_const = getenv("OQRL_HBM_NO_CONST") != nullptr;
        const bool const_single = nb == 1 && recs <= kConstCoefRecs && !no_const;
        // x / n_rows == umulhi(x, ceil(2^32 / n_rows)) for every x < 2^32 / n_rows; n_rows = 1 needs no table per candidate
        const bool const_multi = nb > 1 && !sp.reupload && n_rows > 1 && (long long)n_cands * recs <= kConstCoefRecs &&
                                 (unsigned long long)(cand_off + nb) * (unsigned long long)n_rows < (1ull << 32) && !no_const;
        const size_t rec_bytes = (size_t)recs * 2 * sizeof(float4);
        if (const_single) {
            OQRL_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_hbm_coef, ws + w.coef_off, rec_bytes, 0, cudaMemcpyDeviceToDevice, st));
        } else if (const_multi) {
            // candidate 0 of the batch is represented by its first circuit, candidate j >= 1 by circuit j * n_rows - cand_off
            char *sym = nullptr;
            OQRL_CUDA_CHECK(cudaGetSymbolAddress((void **)&sym, c_hbm_coef));
            OQRL_CUDA_CHECK(cudaMemcpyAsync(sym, ws + w.coef_off, rec_bytes, cudaMemcpyDeviceToDevice, st));
            if (n_cands > 1)
                OQRL_CUDA_CHECK(cudaMemcpy2DAsync(sym + rec_bytes, rec_bytes, ws + w.coef_off + (size_t)(n_rows - cand_off) * rec_bytes,
                                                  (size_t)n_rows * rec_bytes, rec_bytes, (size_t)(n_cands - 1), cudaMemcpyDeviceToDevice, st));
        }
        const bool const_coef = const_single || const_multi;
        a.cand_recs = const_multi ? recs : 0; a.cand_off = cand_off;
        a.rows_magic = n_rows > 1 ? (uint32_t)(((1ull << 32) + (unsigned long long)n_rows - 1) / (unsigned long long)n_rows) : 0u;
        const long long total_tiles = (long long)tiles_per_circ * nb;
        int grid = sm_count;                                   // one persistent CTA per SM (three 64 KiB tile buffers)
        if ((long long)grid > total_tiles) grid = (int)total_tiles;
        for (int p = 0; p < n_pass; ++p) {
            a.pass = p;
            static const int dbg = []() { const char *v = getenv("OQRL_HBM_DEBUG"); return v ? atoi(v) : 0; }();
            a.flags = plan[p].flags | ((dbg & 1) ? PASS_DBG_NOGATES : 0) | ((dbg & 2) ? PASS_DBG_TIMERS : 0) | ((dbg & 4) ? PASS_DBG_NOCOPY : 0) | ((dbg & 8) ? PASS_DBG_NOSPLIT : 0) | ((dbg & 16) ? PASS_DBG_STAGGER : 0); a.n_tile_bits = plan[p].n_tile_bits;
            a.row_piv_packed[0] = a.row_piv_packed[1] = 0;
            for (int j = 0; j < 16; ++j) a.gate_lw[j] = j < plan[p].n_gates ? (uint16_t)(plan[p].gate[j].layer * n + plan[p].gate[j].wire) : 0;
            for (int j = 0; j < kMaxGroups; ++j) {
                const HbmGroup &hg = plan[p].group[j];
                const bool live = j < plan[p].n_groups;
                a.group_g0[j] = live ? (uint8_t)hg.gate[0] : 0;
                a.group_uni[j] = live && const_coef && (hg.rot_lane[0] & hg.rot_lane[1] & hg.rot_lane[2] & hg.rot_lane[3]) == 0xff;
            }
            for (int j = 0; j < kRowBits; ++j) {
                a.row_vec[j] = plan[p].row_vec[j];
                a.row_piv_packed[j >> 2] |= (uint32_t)plan[p].row_piv[j] << (8 * (j & 3));
            }
            {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kHbmThreads); cfg.dynamicSmemBytes = kHbmSmemBytes; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cudaError_t le = cudaLaunchKernelEx(&cfg, hbm_pass_kernel, a);
                if (le != cudaSuccess) { cudaGetLastError(); set_error("launch of hbm_pass_kernel failed: %s", cudaGetErrorString(le)); return OQRL_ERR_CUDA; }
            }
            OQRL_LAUNCH_CHECK("hbm_pass_kernel");
            if (dbg & 2) {
                static int printed = 0;
                unsigned long long h[18];
                cudaStreamSynchronize(st);
                cudaMemcpy(h, ws + w.fault_off + 8, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                cudaMemsetAsync(ws + w.fault_off, 0, 256, st);
                if (printed++ < 80) {
                    const double tiles = (double)total_tiles;
                    fprintf(stderr, "[hbm pass %d gates %d] consumer cyc/tile: setup %.0f wait_tile %.0f gen_rows %.0f gen_sync %.0f groups %.0f prep %.0f barrier %.0f handoff %.0f | producer: prologue %.0f wait_done %.0f store_issue %.0f wait_read %.0f load_issue %.0f drain %.0f\n",
                            p, plan[p].n_gates, h[0] / tiles, h[1] / tiles, h[7] / tiles, h[2] / tiles, h[3] / tiles, h[4] / tiles, h[5] / tiles, h[6] / tiles,
                            h[8] / tiles, h[9] / tiles, h[10] / tiles, h[11] / tiles, h[12] / tiles, h[13] / tiles);
                }
            }
        }
        if (d_q) {
            const int items = nb * sp.n_out;
            hbm_zreduce_kernel<<<items, kZReduceThreads, 0, st>>>(a.zpart, nb, tiles_per_circ, sp.n_out,
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
    double f = 6.0 * dim + 6.0 * (dim / (double)(1 << kB)) * 3.0;   // generation: one complex multiply per amplitude + the row factors
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
