"""The oracle against its pins (CPU only): SURVEY.md Appendix B KATs, evaluator
cross-checks, C port vs numpy, and the committed golden fitness vectors."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import c_oracle, pqc_oracle as po
import make_golden


def test_basis_state_kats(kat):
    s4 = po.Spec(n_out=4)
    w = np.zeros(24, np.float32)
    for case in kat["basis_cases_n4_L2_w0"]:
        x = np.array([case["x"]], np.float64)
        for ev in (po.qvalues_tensor, po.qvalues_dense):
            np.testing.assert_allclose(ev(s4, w, x)[0], case["z"], atol=1e-6)  # pi is rounded to f32 first


def test_seeded_kats(kat, sample):
    s = po.Spec()
    w = np.linspace(-1.2, 1.1, 24).astype(np.float32)
    q = po.qvalues_tensor(s, w, np.array([kat["linspace_case"]["x"]], np.float32))[0]
    np.testing.assert_allclose(q, kat["linspace_case"]["q"], atol=2e-10)
    w0 = np.random.default_rng(0).random(24).astype(np.float32)
    q = po.qvalues_tensor(s, w0, sample["head_obs"][:4])
    for r in range(4):
        np.testing.assert_allclose(q[r], kat["w0_rows"][str(r)], atol=2e-10)
    f = po.fitness(s, w0[None], sample["head_obs"][:16], sample["head_act"][:16], sample["head_rew"][:16],
                   sample["head_next_obs"][:16], sample["head_term"][:16])
    # Appendix B used a float64 0.99; the reference's torch expression rounds (1-d)*0.99 to float32
    # (optim.py:163), which the oracle follows -> agreement to 0.99f-0.99 ~ 1e-8, not 1e-12.
    np.testing.assert_allclose(f[0], kat["w0_fitness_rows_0_15_gamma_0.99"], atol=5e-8)


def test_sel_ranges(kat):
    for key, want in kat["sel_ranges"].items():
        n, L = map(int, key.split(","))
        assert [po.sel_range(n, l) for l in range(L)] == want


def test_dataset_sample_checksums(sample):
    # Appendix C checksums travel with the fixture
    assert str(sample["sha256"]) == "864fffb43c61538d1fa589e37fdffe9bc568da72ece75872b356aa3061d7a3b5"
    assert int(sample["sum_actions"]) == 50672 and int(sample["sum_terminals"]) == 2342
    np.testing.assert_allclose(sample["sum_obs_cols"], [-4076.70924943, -3102.45794688, -2113.70040331, -9705.67856026], atol=1e-6)
    term = sample["sample_term"].astype(bool)
    assert np.all(sample["sample_next_obs"][term] == 0)      # terminal rows have next_obs = 0
    assert np.all(sample["sample_rew"] == 1.0)


spec_strategy = st.builds(
    po.Spec,
    n_qubits=st.integers(1, 6), n_layers=st.integers(0, 4), n_out=st.just(1), n_feat=st.integers(0, 4),
    entangler=st.sampled_from([po.SEL_CNOT, po.CZ_RING]), reupload=st.integers(0, 1),
    feature_map=st.sampled_from([po.FEAT_FIRST_K, po.FEAT_TILED]))


@settings(max_examples=40, deadline=None)
@given(spec=spec_strategy, seed=st.integers(0, 2**31 - 1))
def test_evaluators_agree(spec, seed):
    if spec.feature_map == po.FEAT_FIRST_K and spec.n_feat > spec.n_qubits:
        with pytest.raises(ValueError):
            spec.validate()
        return
    spec = po.Spec(**{**spec.__dict__, "n_out": spec.n_qubits})
    rng = np.random.default_rng(seed)
    w = rng.normal(size=spec.n_params).astype(np.float32)
    x = rng.normal(size=(3, spec.n_feat)).astype(np.float32) * 2
    qt, qd, qc = po.qvalues_tensor(spec, w, x), po.qvalues_dense(spec, w, x), c_oracle.qvalues(spec, w, x)[0]
    np.testing.assert_allclose(qt, qd, atol=1e-12)
    np.testing.assert_allclose(qt, qc, atol=1e-12)
    assert np.all(np.abs(qt) <= 1 + 1e-12)
    norm = np.sum(np.abs(po.simulate(spec, w, x)) ** 2, axis=1)
    np.testing.assert_allclose(norm, 1.0, atol=1e-12)


def test_zero_everything_gives_plus_one():
    for n in (1, 2, 5, 9):
        spec = po.Spec(n_qubits=n, n_layers=3, n_out=n, n_feat=min(n, 4))
        q = po.qvalues_tensor(spec, np.zeros(spec.n_params), np.zeros((1, spec.n_feat)))
        np.testing.assert_allclose(q, 1.0, atol=1e-15)


def test_c_port_matches_numpy_fitness(sample):
    for name in ("c1_ref", "n4_cz_L3", "n6_tiled_reup_L3", "c3_small"):
        spec, params, obs, act, rew, nxt, term = make_golden.golden_inputs(name, sample)
        f_np = po.fitness(spec, params, obs, act, rew, nxt, term)
        f_c = c_oracle.fitness(spec, params, obs, act, rew, nxt, term)
        np.testing.assert_allclose(f_c, f_np, rtol=1e-13)
        f_c1 = c_oracle.fitness(spec, params, obs, act, rew, nxt, term, nthreads=1)
        np.testing.assert_array_equal(f_c, f_c1)     # thread count must not change results


def test_golden_fitness_regression(sample, golden_fitness):
    for name in make_golden.GOLDEN_SPECS:
        spec, params, obs, act, rew, nxt, term = make_golden.golden_inputs(name, sample)
        if spec.n_qubits > 8:
            f = c_oracle.fitness(spec, params, obs, act, rew, nxt, term)
        else:
            f = po.fitness(spec, params, obs, act, rew, nxt, term)
        np.testing.assert_allclose(f, golden_fitness[name + "_fitness"], rtol=1e-12)


def test_canonical_flop_formula():
    # C1/C2 = 2256, C3 (tiled, reupload) = 328960, C4 (L=2, n_feat=4 first_k) = 31.26 M
    assert po.flops_per_eval(po.Spec()) == 2256
    assert po.flops_per_eval(po.Spec(8, 8, 2, 4, 0, 1, po.FEAT_TILED)) == 328960
    assert po.flops_per_eval(po.Spec(8, 8, 2, 4, 0, 1, po.FEAT_FIRST_K)) == 279808
    assert po.flops_per_eval(po.Spec(16, 2, 2, 4)) == 65536 * (24 + 448 + 5)


def test_rotated_readout_identity_reference_circuit():
    """The reference-circuit kernel never applies layer 1: it measures U^+ Z U on the state before it.

    Pins, with the oracle only, the two facts pqc_reg.cu::readout_rotated relies on (n = 4, L = 2, CNOT ring):
    <Z_0>, <Z_1> after the last ring are <Z_2>, <Z_3> before it, and <Z> after a gate U = [[a, b], [-conj b, conj a]]
    is n_z (S0 - S1) + 4 Re(conj(a) b C) with S0/S1 the wire's half norms and C = sum conj(psi_0) psi_1."""
    spec = po.Spec()
    n = spec.n_qubits
    rng = np.random.default_rng(11)
    params = rng.normal(size=spec.n_params).astype(np.float32)
    x = rng.normal(size=(5, 4)).astype(np.float32)
    ref = po.qvalues_tensor(spec, params, x)
    # state before layer 1: run the one-layer circuit (embedding, layer-0 gates, ring of range 1)
    one = po.Spec(n, 1, spec.n_out, spec.n_feat, spec.entangler, spec.reupload, spec.feature_map)
    psi = po.simulate(one, params[: one.n_params], x)
    w = np.asarray(params, dtype=np.float64).reshape(2, n, 3)
    for i, wire in ((0, 2), (1, 3)):
        u = po.rot_matrix(*w[1, wire])
        a, b = u[0, 0], u[0, 1]
        assert np.allclose(u[1, 0], -np.conj(b)) and np.allclose(u[1, 1], np.conj(a))
        v = psi.reshape(psi.shape[0], 1 << wire, 2, 1 << (n - 1 - wire))
        s0 = (np.abs(v[:, :, 0, :]) ** 2).sum(axis=(1, 2))
        s1 = (np.abs(v[:, :, 1, :]) ** 2).sum(axis=(1, 2))
        c = (np.conj(v[:, :, 0, :]) * v[:, :, 1, :]).sum(axis=(1, 2))
        q = (abs(a) ** 2 - abs(b) ** 2) * (s0 - s1) + 4.0 * np.real(np.conj(a) * b * c)
        np.testing.assert_allclose(q, ref[:, i], atol=1e-12)
