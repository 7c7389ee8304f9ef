"""The shared-memory swizzle constants of oqrl_b200/csrc/pqc_blk.cu against the offline search that produced them
(scripts/blk_swizzle_search.py): GF(2)-linear, bijective inside every 2^n block, conflict-free row loads, and the
relabelling scatter conflicts of the search optimum."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(n):
    src = open(os.path.join(ROOT, "oqrl_b200", "csrc", "pqc_blk.cu")).read()
    m = re.search(r"kTable = N == 5 \? (0x[0-9a-f]+)u : \(N == 6 \? (0x[0-9a-f]+)u : (0x[0-9a-f]+)u\)", src)
    assert m, "swizzle constants not found"
    packed = int(m.group(1 if n == 5 else (2 if n == 6 else 3)), 16)
    return [(packed >> (4 * x)) & 7 for x in range(8)]


def _slot(k, tab):
    return k ^ tab[(k >> 4) & 7]


def test_swizzle_is_linear_bijective_and_optimal(capsys):
    spec = importlib.util.spec_from_file_location("blk_swizzle_search", os.path.join(ROOT, "scripts", "blk_swizzle_search.py"))
    search = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(search)                      # runs the search (a few seconds) and prints its table
    best = {}
    for line in capsys.readouterr().out.splitlines():
        m = re.match(r"n=(\d+): current .* -> best \((\d+), (\d+)\)", line)
        if m:
            best[int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
    for n in range(5, 14):
        tab = _table(n)
        dim = 1 << n
        for a in (1, 5, 0x13, 0x37 % dim, dim - 1):
            for b in (2, 0x1c % dim, 0x55 % dim):
                assert _slot(a ^ b, tab) == _slot(a, tab) ^ _slot(b, tab) ^ _slot(0, tab)
        assert _slot(0, tab) == 0
        assert sorted(_slot(k, tab) for k in range(dim)) == list(range(dim))
        cols = [0] + [tab[1 << j] for j in range(3)] + [0]          # index bits 3..7
        ncol = min(5, n - 3)
        assert search.score(n, cols[:ncol]) == best[n], n
