"""numpy execution of the HBM-class pass list exported by oqrl_hbm_plan (CPU only, test infrastructure).

It follows oqrl_b200/csrc/pqc_hbm.cu::hbm_pass_kernel step by step -- tile addressing, register-block
enumeration over the role-0 subspace, generated first pass, CZ signs, read-out masks in physical and in tile
coordinates -- so the host planner (GF(2) layout tracking) is validated against the oracle without a GPU."""
import ctypes

import numpy as np

from oqrl_b200 import _lib

KT, KB, MAX_OUT = 13, 5, 8
MAX_GROUPS = (KT + 3) // 4
PASS_FIRST, PASS_FINAL, PASS_CZ, PASS_CZ_PRE = 1, 2, 4, 8


class Gate(ctypes.Structure):
    _fields_ = [("layer", ctypes.c_int32), ("wire", ctypes.c_int32), ("delta", ctypes.c_uint32),
                ("rho", ctypes.c_uint32), ("r_phys", ctypes.c_uint32), ("d_phys", ctypes.c_uint32)]


class Group(ctypes.Structure):
    _fields_ = [("n_gates", ctypes.c_int32), ("gate", ctypes.c_int32 * 4), ("tvec", ctypes.c_uint32 * KT), ("rot_lane", ctypes.c_uint8 * 4),
                ("delta", ctypes.c_uint32 * 4), ("r_phys", ctypes.c_uint32 * 4)]


class Pass(ctypes.Structure):
    _fields_ = [("flags", ctypes.c_int32), ("n_tile_bits", ctypes.c_int32), ("tile_pos", ctypes.c_uint8 * 32),
                ("row_vec", ctypes.c_uint32 * (KT - KB)), ("n_gates", ctypes.c_int32), ("gate", Gate * KT),
                ("n_groups", ctypes.c_int32), ("group", Group * MAX_GROUPS), ("z_mask", ctypes.c_uint32 * MAX_OUT),
                ("z_tile", ctypes.c_uint32 * MAX_OUT), ("split_phi", ctypes.c_uint32), ("a_col", ctypes.c_uint32 * 32), ("row_piv", ctypes.c_uint8 * (KT - KB)), ("pad2", ctypes.c_uint8 * (16 - (KT - KB)))]


BANK_STATS = []   # distinct 8-byte bank pairs (tile-coordinate bits 0..3) among the first half-warp's 16 lanes, per (pass, group) emulated


def get_plan(spec, keep_state=False):
    L = _lib.probe()
    cs = spec.to_c()
    nb = ctypes.c_uint64()
    _lib.check(L.oqrl_hbm_plan(ctypes.byref(cs), None, 0, ctypes.byref(nb), int(keep_state)))
    assert nb.value % ctypes.sizeof(Pass) == 0, "HbmPass layout drifted from the emulator's mirror"
    n = nb.value // ctypes.sizeof(Pass)
    arr = (Pass * n)()
    _lib.check(L.oqrl_hbm_plan(ctypes.byref(cs), arr, nb.value, ctypes.byref(nb), int(keep_state)))
    return list(arr)


def _parity(v):
    v = np.asarray(v, dtype=np.uint64)
    out = np.zeros(v.shape, dtype=np.uint64)
    for s in range(32):
        out ^= (v >> np.uint64(s)) & np.uint64(1)
    return out.astype(np.int64)


def _insert_zeros(vals, piv):
    out = vals.copy()
    for p in piv:
        p = int(p)
        out = ((out >> p) << (p + 1)) | (out & ((1 << p) - 1))
    return out


def _deposit(vals, pos):
    out = np.zeros_like(vals)
    for i, p in enumerate(pos):
        out |= ((vals >> i) & 1) << int(p)
    return out


def run_plan(spec, plan, coef, vec0):
    """coef [max(L,1)][n] complex (a, b) pairs; vec0 [n][2] complex.  Returns (q[n_out] or None, logical state or None)."""
    n = spec.n_qubits
    state = np.zeros(1 << n, dtype=np.complex128)
    q = None
    for ps in plan:
        ntb = ps.n_tile_bits
        assert ntb == n - KT
        tiles = _deposit(np.arange(1 << ntb, dtype=np.int64), list(ps.tile_pos)[:ntb])
        assert np.array_equal(tiles, _insert_zeros(np.arange(1 << ntb, dtype=np.int64) << KB, list(ps.row_piv)))
        e = np.arange(1 << KT, dtype=np.int64)
        r, lane = e >> KB, e & ((1 << KB) - 1)
        rowaddr = np.zeros_like(e)
        for j in range(KT - KB):
            rowaddr ^= ((r >> j) & 1) * int(ps.row_vec[j])
        zacc = np.zeros(spec.n_out)
        seen = np.zeros(1 << n, dtype=bool)
        for base in tiles:
            addr = (int(base) ^ rowaddr) + lane
            assert not seen[addr].any()
            seen[addr] = True
            if ps.flags & PASS_FIRST:
                tile = np.ones(1 << KT, dtype=np.complex128)
                for pos in range(n):
                    wire = n - 1 - pos
                    tile = tile * vec0[wire][(addr >> pos) & 1]
                if ps.flags & PASS_CZ_PRE:
                    rot = ((addr << 1) | (addr >> (n - 1))) & ((1 << n) - 1)
                    tile = np.where(_parity(addr & rot) == 1, -tile, tile)
            else:
                tile = state[addr].copy()
            for gi in range(ps.n_groups):
                gr = ps.group[gi]
                G = gr.n_gates
                # blocks always hold 16 amplitudes: missing gate slots are padded with role-0 basis vectors
                beta = np.arange(1 << (KT - 4), dtype=np.int64)
                cb = np.zeros_like(beta)
                for j in range(KT - 4):
                    assert int(gr.tvec[j]) != 0
                    cb ^= ((beta >> j) & 1) * int(gr.tvec[j])
                for i in range(G):           # per tile: onto the coset whose element 0 has role 0 for every gate
                    hg = ps.gate[gr.gate[i]]
                    if bin(int(hg.r_phys) & int(base)).count("1") & 1:
                        cb ^= int(hg.delta)
                phi = int(ps.split_phi)     # split tiles: thread bit 7 alone selects the coset, no direction leaves it
                if phi:
                    for j in range(KT - 4):
                        assert _parity(np.array([phi & int(gr.tvec[j])]))[0] == (1 if j == 7 else 0)
                    for i in range(4):
                        assert _parity(np.array([phi & int(gr.delta[i])]))[0] == 0
                tid = beta & 255             # per-lane rotation of the block (bank-conflict avoidance); 256 consumer threads
                for i in range(4):
                    want = (int(ps.gate[gr.gate[i]].delta), int(ps.gate[gr.gate[i]].r_phys)) if i < G else (int(gr.tvec[KT - 4 + i - G]), 0)
                    assert (int(gr.delta[i]), int(gr.r_phys[i])) == want
                rot = []
                for i in range(G):
                    rl = int(gr.rot_lane[i])
                    r = ((tid >> rl) & 1) if rl != 0xff else np.zeros_like(tid)
                    cb = cb ^ (r * int(ps.gate[gr.gate[i]].delta))
                    rot.append(r.astype(bool))
                BANK_STATS.append(len(set((cb[:16] & 15).tolist())))
                idx = np.empty((16, beta.size), dtype=np.int64)
                for m in range(16):
                    c = cb.copy()
                    for i in range(4):
                        if (m >> i) & 1:
                            c ^= int(gr.delta[i])
                    idx[m] = c
                assert np.array_equal(np.sort(idx.ravel()), e), "register blocks must partition the tile"
                v = tile[idx]
                for i in range(G):
                    hg = ps.gate[gr.gate[i]]
                    a, b = coef[hg.layer][hg.wire]
                    role0 = _parity(int(hg.rho) & cb) ^ (bin(int(hg.r_phys) & int(base)).count("1") & 1)
                    assert np.array_equal(role0.astype(bool), rot[i]), "element 0 has role 0 unless the lane starts from the partner"
                    a = np.where(rot[i], np.conj(a), a)
                    b = np.where(rot[i], -np.conj(b), b)
                    pd = 0                                       # d_phys is the direction as a physical address
                    for j in range(KT - KB):
                        pd ^= ((int(hg.delta) >> (KB + j)) & 1) * int(ps.row_vec[j])
                    assert pd ^ (int(hg.delta) & ((1 << KB) - 1)) == int(hg.d_phys)
                    for m in range(16):
                        if not (m >> i) & 1:
                            x0, x1 = v[m].copy(), v[m | (1 << i)].copy()
                            v[m] = a * x0 + b * x1
                            v[m | (1 << i)] = -np.conj(b) * x0 + np.conj(a) * x1
                tile[idx] = v
            if ps.flags & PASS_FINAL:
                p = np.abs(tile) ** 2
                for i in range(spec.n_out):
                    sign = _parity(addr & int(ps.z_mask[i]))
                    in_tile = _parity(e & int(ps.z_tile[i])) ^ (bin(int(ps.z_mask[i]) & int(base)).count("1") & 1)
                    assert np.array_equal(sign, in_tile), "z_tile must be z_mask restricted to the tile"
                    zacc[i] += np.sum(np.where(sign == 1, -p, p))
            else:
                if ps.flags & PASS_CZ:
                    rot = ((addr << 1) | (addr >> (n - 1))) & ((1 << n) - 1)
                    tile = np.where(_parity(addr & rot) == 1, -tile, tile)
                state[addr] = tile
        assert seen.all(), "tiles must cover the state"
        if ps.flags & PASS_FINAL:
            q = zacc
    logical = None
    if not (plan[-1].flags & PASS_FINAL):
        k = np.arange(1 << n, dtype=np.int64)
        x = np.zeros_like(k)
        for p in range(n):
            x ^= ((k >> p) & 1) * int(plan[-1].a_col[p])
        logical = state[x]
    return q, logical
