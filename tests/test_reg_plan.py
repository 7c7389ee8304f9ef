"""Host logic of the 1..4-qubit fitness kernel (pqc_reg.cu::plan_units), checked on the CPU through the probe library:
the unit list dealt to the persistent warps covers every (candidate, 256-transition segment) exactly once, the static
deal is balanced, and split candidates stay within the partial-sum scratch."""
import ctypes

import numpy as np
import pytest

from oqrl_b200 import _lib

SEG = 256


def plan(n_cand, n_trans, n_warps, can_split=1, cap=1 << 40):
    out = (ctypes.c_int32 * 7)()
    _lib.check(_lib.probe().oqrl_reg_plan(n_cand, n_trans, n_warps, can_split, cap, out))
    return dict(zip(("n_full", "n_big", "big_per_cand", "seg_big", "seg_rest", "n_units", "n_segs"), out))


def decode(p, u):
    """(candidate, first segment, segment count) of unit u -- the kernel's decode()."""
    if u < p["n_full"]:
        return u, 0, p["n_segs"]
    v = u - p["n_full"]
    if v < p["n_big"]:
        c = v // p["big_per_cand"]
        return p["n_full"] + c, (v - c * p["big_per_cand"]) * p["seg_big"], p["seg_big"]
    return p["n_full"] + (v - p["n_big"]), p["big_per_cand"] * p["seg_big"], p["seg_rest"]


CASES = [(4096, 1024, 2368), (2368, 1024, 2368), (4736, 1024, 2368), (25, 16, 2368), (1, 1, 2368), (1, 100000, 2368),
         (3, 4000, 2368), (2400, 257, 2368), (2500, 1000, 2368), (5000, 300, 2368), (40, 1000, 2368), (700, 300, 592),
         (2369, 2752, 2368), (10000, 255, 2368), (10000, 256, 2368), (10000, 257, 2368), (37, 513, 16), (16, 513, 16)]


@pytest.mark.parametrize("n_cand,n_trans,n_warps", CASES)
def test_units_cover_every_segment_once(n_cand, n_trans, n_warps):
    p = plan(n_cand, n_trans, n_warps)
    n_segs = -(-n_trans // SEG)
    assert p["n_segs"] == n_segs
    seen = np.zeros((n_cand, n_segs), dtype=np.int32)
    for u in range(p["n_units"]):
        c, s0, ns = decode(p, u)
        assert 0 <= c < n_cand and ns >= 1 and s0 + ns <= n_segs
        seen[c, s0:s0 + ns] += 1
    assert (seen == 1).all()
    # whole candidates come in full rounds of the warps; only a last round that fills less than half of them is cut
    rest = n_cand % n_warps
    if p["n_full"] != n_cand:
        assert p["n_full"] == n_cand - rest and 2 * rest < n_warps and n_segs > 1
        # every warp's share of the cut round: at most seg_big segments from one unit, within one segment of even
        assert p["seg_big"] == -(-rest * n_segs // n_warps)
        per_cand = p["big_per_cand"] + (1 if p["seg_rest"] else 0)
        assert p["n_units"] == p["n_full"] + rest * per_cand
    # static deal: slot s runs units s, s + W, ...; the busiest warp is within one unit of the average
    work = np.zeros(n_warps)
    for u in range(p["n_units"]):
        work[u % n_warps] += decode(p, u)[2]
    assert work.max() - work.sum() / n_warps <= n_segs


def test_no_scratch_or_small_scratch_keeps_candidates_whole():
    assert plan(100, 1024, 2368, can_split=0)["n_units"] == 100
    assert plan(100, 1024, 2368, can_split=0)["n_full"] == 100
    assert plan(100, 1024, 2368, cap=399)["n_full"] == 100          # 100 candidates x 4 segments do not fit 399 doubles
    assert plan(100, 1024, 2368, cap=400)["n_full"] == 0
    assert plan(100, 200, 2368)["n_full"] == 100                    # a single segment cannot be cut
