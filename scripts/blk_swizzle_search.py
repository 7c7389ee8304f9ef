"""Offline search of the shared-memory swizzle of pqc_blk.cu.

slot(k) = k ^ T[(k >> 3) & 31]: the low three bits (the 16-byte bank group of a record) are XORed with a GF(2)-linear
function of index bits 3..7.  Constraint: the row-wise load (eight lanes on index bits 4, 5, 6) stays conflict-free.
Objective: bank conflicts of the relabelling scatter (eight lanes on index bits 0, 1, 2, images under the CNOT ring of
every range) summed over the ranges, then those of the first scatter.  Prints the per-qubit-count column table that is
pasted into pqc_blk.cu (kSwizzleCols).
"""
import itertools


def sel_perm(k, n, r):
    for i in range(n):
        t = (i + r) % n
        if (k >> (n - 1 - i)) & 1:
            k ^= 1 << (n - 1 - t)
    return k


def rank3(vs):
    basis = []
    for v in vs:
        for b in basis:
            v = min(v, v ^ b)
        if v:
            basis.append(v)
    return len(basis)


def bank(k, cols):
    b = k & 7
    for j, c in enumerate(cols):
        if (k >> (3 + j)) & 1:
            b ^= c
    return b


def score(n, cols):
    hi = [b for b in (4, 5, 6) if b < n]
    if rank3([bank(1 << b, cols) for b in hi]) < len(hi):
        return None
    lanes = min(3, n - 4)                      # lane bits inside one transition's group
    tot = 0
    for r in range(1, n):
        tot += 2 ** (lanes - rank3([bank(sel_perm(1 << b, n, r), cols) for b in range(lanes)]))
    first = 2 ** (len(hi) - rank3([bank(sel_perm(1 << b, n, 1), cols) for b in hi]))
    return tot, first


for n in range(5, 14):
    ncol = min(5, n - 3)
    best = None
    for cols in itertools.product(range(8), repeat=ncol):
        s = score(n, cols)
        if s is not None and (best is None or s < best[0]):
            best = (s, cols)
    cur = score(n, ([0, 1, 2, 4, 0])[:ncol])
    cols = list(best[1]) + [0] * (5 - ncol)
    print(f"n={n}: current {cur} -> best {best[0]}  cols {{{', '.join(map(str, cols))}}}")
