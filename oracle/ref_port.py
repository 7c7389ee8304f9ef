"""Per-sample, per-candidate port of the reference's control flow -- TEST INFRASTRUCTURE.

bench.py times this as the "reference-structured" CPU row:
the same loop nest as /root/reference/src/optim.py:204 (one candidate at a
time) -> agent.py:51 (one sample at a time) -> one 2^n-amplitude complex128
state updated gate by gate with small tensor contractions, which is how
PennyLane's default.qubit applies operations.  It leaves OUT PennyLane's tape
construction and template expansion (not reproducible here), so it flatters the
reference.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np

from . import pqc_oracle as po


def _circuit_one_sample(spec, w, x):
    n = spec.n_qubits
    psi = np.zeros([2] * n, dtype=np.complex128)
    psi[(0,) * n] = 1.0

    def gate(m, wire):
        nonlocal psi
        psi = np.moveaxis(np.tensordot(m, psi, axes=([1], [wire])), 0, wire)

    def embed():
        for wire, f in po.embed_wires(spec):
            gate(po.ry_matrix(np.float64(x[f])).astype(np.complex128), wire)

    if not spec.reupload or spec.n_layers == 0:
        embed()
    for l in range(spec.n_layers):
        if spec.reupload:
            embed()
        for i in range(n):
            gate(po.rot_matrix(*w[l, i]), i)
        if n > 1:
            flat = psi.reshape(1, -1)
            if spec.entangler == po.SEL_CNOT:
                r = po.sel_range(n, l)
                for i in range(n):
                    flat = po._apply_cnot(flat, i, (i + r) % n, n)
            else:
                for i in range(n):
                    flat = po._apply_cz(flat, i, (i + 1) % n, n)
            psi = flat.reshape([2] * n)
    return po.expval_z(psi.reshape(1, -1), n, spec.n_out)[0]


def forward_per_sample(spec, params, xs):
    """agent.py:48-61: python loop over rows, one circuit each, stacked."""
    w = np.asarray(params, dtype=np.float32).astype(np.float64).reshape(spec.n_layers, spec.n_qubits, 3)
    return np.stack([_circuit_one_sample(spec, w, np.asarray(x, dtype=np.float32)) for x in xs])


def fitness_per_candidate(spec, population, obs, act, rew, next_obs, term, gamma=0.99):
    """optim.py:204: [evaluate(ind) for ind in population], each 2*T per-sample circuit calls."""
    out = []
    for ind in population:
        q = forward_per_sample(spec, ind, obs)[None]
        qn = forward_per_sample(spec, ind, next_obs)[None]
        out.append(po.td_fitness_from_q(q, qn, act, rew, term, gamma)[0])
    return np.array(out)
