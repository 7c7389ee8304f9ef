"""Host planner of the HBM-resident class (GF(2) layout tracking, tiles, register blocks) executed in
numpy and checked against the oracle (CPU only)."""
import numpy as np
import pytest

from conftest import to_product_spec
from hbm_plan_emulator import PASS_FINAL, PASS_FIRST, get_plan, run_plan
from oracle import pqc_oracle as po


def _tables(spec, w, x):
    """What hbm_prep_kernel computes: fused SU(2) coefficients per (layer, wire) and layer-0 product vectors."""
    n, L = spec.n_qubits, spec.n_layers
    emb = dict(po.embed_wires(spec))
    wl = np.asarray(w, np.float64).reshape(L, n, 3) if L else np.zeros((0, n, 3))
    coef = [[None] * n for _ in range(max(L, 1))]
    vec0 = []
    for wire in range(n):
        c, s = (np.cos(x[emb[wire]] / 2), np.sin(x[emb[wire]] / 2)) if wire in emb else (1.0, 0.0)
        m0 = po.rot_matrix(*wl[0, wire]) if L else np.eye(2)
        vec0.append(m0 @ np.array([c, s]))
        for l in range(1, L):
            m = po.rot_matrix(*wl[l, wire])
            if spec.reupload and wire in emb:
                m = m @ po.ry_matrix(x[emb[wire]])
            coef[l][wire] = (m[0, 0], m[0, 1])
    return coef, vec0


CASES = [po.Spec(15, 2, 2, 4), po.Spec(15, 3, 3, 4, po.SEL_CNOT, 1, po.FEAT_TILED), po.Spec(16, 2, 2, 4, po.CZ_RING, 0, po.FEAT_TILED),
         po.Spec(15, 1, 2, 4), po.Spec(15, 0, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), po.Spec(15, 4, 2, 4, po.CZ_RING, 1, po.FEAT_TILED),
         po.Spec(17, 3, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), po.Spec(14, 3, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
         po.Spec(14, 2, 2, 4, po.CZ_RING, 0, po.FEAT_FIRST_K)]


@pytest.mark.parametrize("spec", CASES, ids=lambda s: f"n{s.n_qubits}L{s.n_layers}e{s.entangler}r{s.reupload}")
def test_plan_matches_oracle(spec):
    rng = np.random.default_rng(spec.n_qubits * 100 + spec.n_layers)
    w = rng.normal(size=spec.n_params).astype(np.float32)
    x = rng.normal(size=spec.n_feat).astype(np.float32).astype(np.float64)
    coef, vec0 = _tables(spec, w, x)
    ps = to_product_spec(spec)
    plan = get_plan(ps)
    assert plan[0].flags & PASS_FIRST and plan[-1].flags & PASS_FINAL
    q, _ = run_plan(spec, plan, coef, vec0)
    ref = po.qvalues_tensor(spec, w, x[None])[0]
    np.testing.assert_allclose(q, ref, atol=1e-10)
    # statevector hook: same plan without the FINAL flag, un-permuted through the tracked layout
    plan2 = get_plan(ps, keep_state=True)
    assert not plan2[-1].flags & PASS_FINAL
    _, logical = run_plan(spec, plan2, coef, vec0)
    np.testing.assert_allclose(logical, po.simulate(spec, w, x[None])[0], atol=1e-10)


def test_sweep_counts():
    """Passes per circuit (what hbm_sweeps_per_eval reports) stay far below the gate-by-gate count."""
    for n, L in ((16, 2), (24, 2), (24, 8), (20, 4)):
        plan = get_plan(to_product_spec(po.Spec(n, L, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED)))
        gate_by_gate = n * L
        assert len(plan) <= (L - 1) * -(-n // 8) + 1 < gate_by_gate
