"""numpy execution of the HBM-class pass list exported by oqrl_hbm_plan (CPU only, test infrastructure).

It follows oqrl_b200/csrc/pqc_hbm.cu::hbm_pass_kernel step by step -- tile addressing, register-block
enumeration, per-block role swaps, generated first pass, CZ signs, read-out masks -- so the host planner
(GF(2) layout tracking) is validated against the oracle without a GPU."""
import ctypes

import numpy as np

from oqrl_b200 import _lib

KT, KB, MAX_OUT = 13, 4, 8
SLOTS = 12
PASS_FIRST, PASS_FINAL, PASS_CZ, PASS_CZ_PRE = 1, 2, 4, 8


class Gate(ctypes.Structure):
    _fields_ = [("layer", ctypes.c_int32), ("wire", ctypes.c_int32), ("rho", ctypes.c_uint32), ("r_phys", ctypes.c_uint32)]


class Pass(ctypes.Structure):
    _fields_ = [("flags", ctypes.c_int32), ("n_tile_bits", ctypes.c_int32), ("tile_pos", ctypes.c_uint8 * 32),
                ("row_vec", ctypes.c_uint32 * (KT - KB)), ("row_piv", ctypes.c_uint8 * (KT - KB)),
                ("pad", ctypes.c_uint8 * (16 - (KT - KB))), ("n_gates", ctypes.c_int32), ("n_groups", ctypes.c_int32),
                ("gate", Gate * SLOTS), ("t_col", ctypes.c_uint16 * KT), ("pad2", ctypes.c_uint16 * 3),
                ("z_mask", ctypes.c_uint32 * MAX_OUT), ("a_col", ctypes.c_uint32 * 32)]


BANK_STATS = []   # distinct 8-byte bank slots among the first warp's lanes, per (pass, group) emulated


def get_plan(spec, keep_state=False):
    L = _lib.lib()
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
    """coef [max(L,1)][n] complex (a, b) pairs; vec0 [n][2] complex.  Returns (q[n_out] or None, logical state or None).

    Mirrors hbm_pass_kernel: rows -> shared-memory coordinates through t_col, three register-block rounds with
    coordinate bit 0 pairing blocks A/B, lane rotation in group 0, role swaps from rho / r_phys."""
    n = spec.n_qubits
    state = np.zeros(1 << n, dtype=np.complex128)
    q = None
    for ps in plan:
        ntb = ps.n_tile_bits
        assert ntb == n - KT
        tiles = _deposit(np.arange(1 << ntb, dtype=np.int64), list(ps.tile_pos)[:ntb])
        assert np.array_equal(tiles, _insert_zeros(np.arange(1 << ntb, dtype=np.int64) << KB, list(ps.row_piv)))
        y = np.arange(1 << KT, dtype=np.int64)                 # tile-relative index = row << KB | low
        r, lane = y >> KB, y & ((1 << KB) - 1)
        rowaddr = np.zeros_like(y)
        for j in range(KT - KB):
            rowaddr ^= ((r >> j) & 1) * int(ps.row_vec[j])
        coord = np.zeros_like(y)                                # shared-memory coordinate of y
        for j in range(KT):
            coord ^= ((y >> j) & 1) * int(ps.t_col[j])
        assert np.array_equal(np.sort(coord), y), "t_col must be a bijection of the tile"
        zacc = np.zeros(spec.n_out)
        seen = np.zeros(1 << n, dtype=bool)
        for base in tiles:
            addr = (int(base) ^ rowaddr) + lane
            assert not seen[addr].any()
            seen[addr] = True
            tile = np.zeros(1 << KT, dtype=np.complex128)       # indexed by shared-memory coordinate
            if ps.flags & PASS_FIRST:
                vals = np.ones(1 << KT, dtype=np.complex128)
                for pos in range(n):
                    wire = n - 1 - pos
                    vals = vals * np.asarray(vec0[wire])[(addr >> pos) & 1]
                if ps.flags & PASS_CZ_PRE:
                    rot = ((addr << 1) | (addr >> (n - 1))) & ((1 << n) - 1)
                    vals = np.where(_parity(addr & rot) == 1, -vals, vals)
                tile[coord] = vals
            else:
                tile[coord] = state[addr]
            t = np.arange(256, dtype=np.int64)
            for gi in range(3):
                if not (ps.n_groups >> gi) & 1:
                    continue
                sh = 1 + 4 * gi
                ca = ((t & ((1 << (sh - 1)) - 1)) << 1) | ((t >> (sh - 1)) << (sh + 4))
                rotv = (t & 15) if gi == 0 else np.zeros_like(t)
                idx = np.empty((2, 16, 256), dtype=np.int64)     # [A/B][register element m][thread]
                for m in range(16):
                    el = (m ^ rotv)
                    idx[0, m] = ca | (el << sh)
                    idx[1, m] = idx[0, m] | 1
                assert np.array_equal(np.sort(idx.ravel()), y), "register blocks must partition the tile"
                BANK_STATS.append(len(set(((idx[0, 0, :32] >> 1) & 15).tolist())) * 2)   # 64-bit accesses: 16 slots x 2 wavefronts
                v = tile[idx]
                c0 = ca | (rotv << sh)
                for i in range(4):
                    hg = ps.gate[4 * gi + i]
                    if hg.wire < 0:
                        continue                                  # identity slot
                    a, b = coef[hg.layer][hg.wire]
                    assert int(hg.rho) == 1 << (sh + i), "fillers must be role-neutral (kernel relies on it)"
                    rho0 = bin(int(hg.r_phys) & int(base)).count("1") & 1
                    sa = (_parity(int(hg.rho) & c0) ^ rho0).astype(bool)
                    sb = sa ^ bool(int(hg.rho) & 1)
                    for half, sflag in ((0, sa), (1, sb)):
                        ga = np.where(sflag, np.conj(a), a)
                        gb = np.where(sflag, -np.conj(b), b)
                        for m in range(16):
                            if not (m >> i) & 1:
                                x0, x1 = v[half, m].copy(), v[half, m | (1 << i)].copy()
                                v[half, m] = ga * x0 + gb * x1
                                v[half, m | (1 << i)] = -np.conj(gb) * x0 + np.conj(ga) * x1
                tile[idx] = v
            out = tile[coord]
            if ps.flags & PASS_FINAL:
                p = np.abs(out) ** 2
                for i in range(spec.n_out):
                    zacc[i] += np.sum(np.where(_parity(addr & int(ps.z_mask[i])) == 1, -p, p))
            else:
                if ps.flags & PASS_CZ:
                    rot = ((addr << 1) | (addr >> (n - 1))) & ((1 << n) - 1)
                    out = np.where(_parity(addr & rot) == 1, -out, out)
                state[addr] = out
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
