"""Parity of the CUDA path against the oracle, through the C ABI (B200 only: -m gpu)."""
import os

import numpy as np
import pytest
import torch

import make_golden
from conftest import assert_same_ranking, rel_err, to_product_spec
from oracle import c_oracle, pqc_oracle as po

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north_star: Q-values and fitness within 1e-5 relative in fp32 (floor max(|ref|,1))


@pytest.fixture(scope="module")
def eng():
    import oqrl_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    path = oqrl_b200._build.LIB_PATH
    assert os.path.exists(path), "liboqrl_b200.so must be built in-tree"
    return oqrl_b200.engine


FAMILY_CASES = [n for n in make_golden.GOLDEN_SPECS]


@pytest.mark.parametrize("name", FAMILY_CASES)
def test_golden_cases_qvalues_and_fitness(name, eng, sample, golden_fitness):
    spec, params, obs, act, rew, nxt, term = make_golden.golden_inputs(name, sample)
    ps = to_product_spec(spec)
    q = eng.qvalues(ps, params, obs).cpu().numpy()
    assert q.shape == (params.shape[0], obs.shape[0], spec.n_out)
    assert rel_err(q, golden_fitness[name + "_q"]).max() < TOL
    f = eng.population_fitness_raw(ps, params, obs, act, rew, nxt, term).cpu().numpy()
    ref = golden_fitness[name + "_fitness"]
    assert rel_err(f, ref).max() < TOL
    assert_same_ranking(f, ref)


def test_appendix_b_kats_through_vqc(eng, kat, sample):
    from oqrl_b200 import VQC
    m = VQC(4, 4, n_layers=2, rng=np.random.default_rng(0))
    m.params.data.zero_()
    for case in kat["basis_cases_n4_L2_w0"]:
        q = m(torch.tensor([case["x"]], dtype=torch.float32))
        assert q.dtype == torch.float64 and q.shape == (1, 4)
        np.testing.assert_allclose(q.cpu().numpy()[0], case["z"], atol=2e-6)
    m2 = VQC(4, 2, n_layers=2, rng=np.random.default_rng(0))       # w0 = first draw of seed 0
    q = m2(torch.tensor(sample["head_obs"][:4])).cpu().numpy()
    for r in range(4):
        np.testing.assert_allclose(q[r], kat["w0_rows"][str(r)], atol=TOL)
    assert int(m2(torch.tensor(sample["head_obs"][:1])).argmax().item()) == 0     # agent.py:116 usage
    m2.params.data.copy_(torch.tensor(np.linspace(-1.2, 1.1, 24).astype(np.float32)))
    q = m2(torch.tensor([kat["linspace_case"]["x"]], dtype=torch.float32)).cpu().numpy()[0]
    np.testing.assert_allclose(q, kat["linspace_case"]["q"], atol=TOL)
    f = eng.population_fitness_raw(m2.spec, np.random.default_rng(0).random(24).astype(np.float32)[None],
                                   sample["head_obs"][:16], sample["head_act"][:16], sample["head_rew"][:16],
                                   sample["head_next_obs"][:16], sample["head_term"][:16]).cpu().numpy()
    assert abs(f[0] - kat["w0_fitness_rows_0_15_gamma_0.99"]) < TOL


def _c2_inputs(sample, P, T, seed=0):
    rng = np.random.default_rng(seed)
    base = rng.random(24).astype(np.float32)
    params = (base[None] + 0.1 * rng.normal(size=(P, 24))).astype(np.float32)      # agent.py:30 + optim.py:58-60
    rows = rng.choice(sample["sample_obs"].shape[0], T, replace=False)
    return params, rows


def test_dataset_residency_and_index_path(eng, sample):
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    assert len(ds) == 4096 and ds.n_feat == 4
    params, rows = _c2_inputs(sample, 64, 256)
    spec = to_product_spec(po.Spec())
    f_idx = eng.population_fitness(spec, params, ds, rows).cpu().numpy()
    f_raw = eng.population_fitness_raw(spec, params, sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                                       sample["sample_next_obs"][rows], sample["sample_term"][rows]).cpu().numpy()
    np.testing.assert_array_equal(f_idx, f_raw)         # same kernel, same order: bit-identical
    obs, nxt, act, rew, term = ds.gather(rows)
    np.testing.assert_array_equal(obs.cpu().numpy(), sample["sample_obs"][rows])
    np.testing.assert_array_equal(nxt.cpu().numpy(), sample["sample_next_obs"][rows])
    np.testing.assert_array_equal(act.cpu().numpy(), sample["sample_act"][rows])
    np.testing.assert_array_equal(term.cpu().numpy(), sample["sample_term"][rows].astype(np.float32))
    with pytest.raises(IndexError):
        ds.gather(np.array([0, 4096]))
    hp = torch.tensor(params).pin_memory()
    hi = torch.tensor(rows.astype(np.int32)).pin_memory()
    ho = torch.empty(64, dtype=torch.float32).pin_memory()
    eng.population_fitness_host(spec, hp, ds, hi, ho)       # pinned: parameters and fitness used in place over PCIe
    np.testing.assert_array_equal(ho.numpy(), f_idx)
    ho2 = torch.empty(64, dtype=torch.float32)              # pageable: staged copies
    eng.population_fitness_host(spec, torch.tensor(params), ds, torch.tensor(rows.astype(np.int32)), ho2)
    np.testing.assert_array_equal(ho2.numpy(), f_idx)
    spec8 = to_product_spec(po.Spec(8, 2, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED))
    p8 = np.random.default_rng(2).normal(size=(9, spec8.n_params)).astype(np.float32)
    want8 = eng.population_fitness(spec8, p8, ds, rows).cpu().numpy()
    ho8 = torch.empty(9, dtype=torch.float32).pin_memory()
    eng.population_fitness_host(spec8, torch.tensor(p8).pin_memory(), ds, hi, ho8)
    np.testing.assert_array_equal(ho8.numpy(), want8)


def test_c2_full_size_parity_ranking_determinism(eng, sample):
    """BASELINE config 2 at full size: 4096 candidates x 1024 transitions."""
    P, T = 4096, 1024
    params, rows = _c2_inputs(sample, P, T)
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    spec = to_product_spec(po.Spec())
    before = eng.launch_count()
    f1 = eng.population_fitness(spec, params, ds, rows)
    assert eng.launch_count() - before == 1                # one fused launch per generation
    f2 = eng.population_fitness(spec, params, ds, rows)
    assert torch.equal(f1, f2)                             # run-to-run bit-deterministic
    f = f1.cpu().numpy()
    ref = c_oracle.fitness(po.Spec(), params, sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                           sample["sample_next_obs"][rows], sample["sample_term"][rows])
    assert rel_err(f, ref).max() < TOL
    swaps = assert_same_ranking(f, ref)
    print(f"config 2: all {P} candidates ranked against the oracle, {swaps} tolerated tie swap(s), max rel err {rel_err(f, ref).max():.2e}")
    # size-independent properties: permuting transitions leaves the mean unchanged up to fp32 summation order,
    # permuting candidates permutes the output exactly
    perm = np.random.default_rng(1).permutation(P)
    fp = eng.population_fitness(spec, params[perm], ds, rows).cpu().numpy()
    np.testing.assert_array_equal(fp, f[perm])
    ft = eng.population_fitness(spec, params[:256], ds, rows[::-1].copy()).cpu().numpy()
    assert rel_err(ft, f[:256]).max() < 1e-6


def test_whole_and_split_work_items_agree(eng, sample):
    """Populations larger than one wave of resident warps mix whole-candidate and split work items (pqc_reg.cu):
    ragged transition counts, oracle parity, and the same bits for a candidate wherever it sits."""
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    spec = to_product_spec(po.Spec())
    for P, T in ((2400, 257), (2500, 1000), (5000, 300)):
        params, rows = _c2_inputs(sample, P, T, seed=P + T)
        f = eng.population_fitness(spec, params, ds, rows).cpu().numpy()
        ref = c_oracle.fitness(po.Spec(), params, sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                               sample["sample_next_obs"][rows], sample["sample_term"][rows])
        assert rel_err(f, ref).max() < TOL, (P, T)
        rev = eng.population_fitness(spec, params[::-1].copy(), ds, rows).cpu().numpy()
        np.testing.assert_array_equal(rev, f[::-1])        # first <-> last: whole items become split ones and back
        small = eng.population_fitness(spec, params[:40], ds, rows).cpu().numpy()    # all split (few candidates)
        np.testing.assert_array_equal(small, f[:40])


def test_staged_batch_and_direct_loads_agree(eng, sample):
    """The 1..4-qubit fitness kernel stages the generation's batch in shared memory when it fits and gathers it with
    direct loads otherwise (pqc_reg.cu): same arithmetic, so the same bits -- for the specialised reference agent and
    the run-time-loop kernels, whole and split work units, and a batch just past the staging capacity."""
    import dataclasses
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    rng = np.random.default_rng(23)
    for spec, P, T in ((po.Spec(), 2500, 1000), (po.Spec(), 37, 257), (po.Spec(4, 3, 2, 4, po.CZ_RING), 300, 512),
                       (po.Spec(3, 4, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), 200, 700), (po.Spec(2, 2, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), 50, 100)):
        ps = to_product_spec(spec)
        params = rng.normal(size=(P, spec.n_params)).astype(np.float32)
        rows = rng.integers(0, sample["sample_obs"].shape[0], T).astype(np.int32)
        staged = eng.population_fitness(ps, params, ds, rows).cpu().numpy()
        direct = eng.population_fitness(dataclasses.replace(ps, direct_loads=1), params, ds, rows).cpu().numpy()
        np.testing.assert_array_equal(staged, direct, err_msg=f"n={spec.n_qubits} L={spec.n_layers} P={P} T={T}")
        ref = c_oracle.fitness(spec, params[:64], sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                               sample["sample_next_obs"][rows], sample["sample_term"][rows])
        assert rel_err(staged[:64], ref).max() < TOL
    # the largest batch that still fits beside the tables of the 2-layer agent (86 blocks of 32 transitions + the spare
    # one), the first that does not (the launcher falls back to direct loads by itself), and a much larger one
    spec = po.Spec()
    params = rng.normal(size=(64, spec.n_params)).astype(np.float32)
    for T in (2752, 2753, 3000, 1, 31, 33):
        rows = rng.integers(0, sample["sample_obs"].shape[0], T).astype(np.int32)
        f = eng.population_fitness(to_product_spec(spec), params, ds, rows).cpu().numpy()
        g = eng.population_fitness(dataclasses.replace(to_product_spec(spec), direct_loads=1), params, ds, rows).cpu().numpy()
        np.testing.assert_array_equal(f, g, err_msg=f"T={T}")
        ref = c_oracle.fitness(spec, params, sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                               sample["sample_next_obs"][rows], sample["sample_term"][rows])
        assert rel_err(f, ref).max() < TOL, T


def test_fitness_bits_do_not_depend_on_population_size(eng, sample):
    """A candidate's fitness is the same float whether it is evaluated alone, in a small block or inside a large
    population (different chunking of the transitions): what makes population sharding over GPUs reproducible."""
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    rng = np.random.default_rng(17)
    rows = rng.integers(0, sample["sample_obs"].shape[0], 1000).astype(np.int32)
    for spec, P in ((po.Spec(), 3000), (po.Spec(6, 2, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), 700),
                    (po.Spec(8, 3, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), 700), (po.Spec(10, 2, 2, 4, po.CZ_RING, 0, po.FEAT_TILED), 160)):
        ps = to_product_spec(spec)
        params = rng.normal(size=(P, spec.n_params)).astype(np.float32)
        full = eng.population_fitness(ps, params, ds, rows).cpu().numpy()
        for k in (1, 7, 90):
            part = eng.population_fitness(ps, params[:k], ds, rows).cpu().numpy()
            np.testing.assert_array_equal(part, full[:k], err_msg=f"n={spec.n_qubits} k={k}")


def test_pruned_and_unpruned_last_layer_agree(eng, sample):
    """The last-layer light-cone pruning is exact: both variants match the oracle (and each other to fp32 noise)."""
    import dataclasses
    for name in ("c1_ref", "n4_L5", "n4_cz_L3", "n4_reup_L3", "n3_L4"):
        spec, params, obs, act, rew, nxt, term = make_golden.golden_inputs(name, sample)
        ps = to_product_spec(spec)
        ref = po.fitness(spec, params, obs, act, rew, nxt, term)
        f1 = eng.population_fitness_raw(ps, params, obs, act, rew, nxt, term).cpu().numpy()
        f2 = eng.population_fitness_raw(dataclasses.replace(ps, no_prune=1), params, obs, act, rew, nxt, term).cpu().numpy()
        assert rel_err(f1, ref).max() < TOL and rel_err(f2, ref).max() < TOL, name
    # all four outputs requested: nothing can be pruned, same kernel family
    spec4 = po.Spec(4, 2, 4, 4)
    params, rows = _c2_inputs(sample, 8, 64, seed=9)
    q = eng.qvalues(to_product_spec(spec4), params, sample["sample_obs"][rows]).cpu().numpy()
    ref = np.stack([po.qvalues_tensor(spec4, p, sample["sample_obs"][rows]) for p in params])
    assert rel_err(q, ref).max() < TOL


def test_transition_split_path_matches_single_block_path(eng, sample):
    """Few candidates x many transitions takes the blockIdx.y-split + finalize path."""
    params, rows = _c2_inputs(sample, 3, 4000, seed=5)
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    spec = to_product_spec(po.Spec())
    f = eng.population_fitness(spec, params, ds, rows).cpu().numpy()
    ref = c_oracle.fitness(po.Spec(), params, sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                           sample["sample_next_obs"][rows], sample["sample_term"][rows])
    assert rel_err(f, ref).max() < TOL


def test_every_qubit_count_dispatches(eng, sample):
    """n = 1..18 all reach a kernel family (register / shared memory / HBM) and match the oracle."""
    rng = np.random.default_rng(99)
    for n in range(1, 19):
        nf = min(n, 4)
        spec = po.Spec(n, 2, min(n, 2), nf, po.SEL_CNOT, 0, po.FEAT_TILED if n > 4 else po.FEAT_FIRST_K)
        params = rng.normal(size=(2, spec.n_params)).astype(np.float32)
        x = np.ascontiguousarray(sample["sample_obs"][:3, :nf])
        q = eng.qvalues(to_product_spec(spec), params, x).cpu().numpy()
        ref = c_oracle.qvalues(spec, params, x)
        assert rel_err(q, ref).max() < TOL, n


BLK_CASES = [  # (n, L, n_out, n_feat, entangler, reupload, feature_map)
    (8, 0, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), (8, 1, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), (8, 2, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
    (8, 3, 4, 4, po.SEL_CNOT, 0, po.FEAT_FIRST_K), (8, 8, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), (8, 9, 3, 8, po.SEL_CNOT, 1, po.FEAT_FIRST_K),
    (8, 2, 1, 4, po.CZ_RING, 1, po.FEAT_TILED), (8, 5, 4, 3, po.CZ_RING, 0, po.FEAT_FIRST_K), (8, 15, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
    (5, 2, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), (5, 3, 5, 4, po.CZ_RING, 0, po.FEAT_FIRST_K), (6, 4, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
    (7, 5, 3, 4, po.SEL_CNOT, 0, po.FEAT_TILED), (7, 2, 2, 7, po.CZ_RING, 1, po.FEAT_FIRST_K), (9, 3, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
    (9, 2, 8, 9, po.CZ_RING, 1, po.FEAT_FIRST_K), (10, 4, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), (10, 1, 3, 4, po.CZ_RING, 0, po.FEAT_TILED),
    (11, 3, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), (11, 2, 4, 6, po.CZ_RING, 1, po.FEAT_TILED), (12, 6, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
    (12, 2, 5, 4, po.CZ_RING, 0, po.FEAT_FIRST_K), (12, 0, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), (13, 3, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
    (13, 2, 3, 5, po.CZ_RING, 1, po.FEAT_TILED), (13, 1, 2, 4, po.SEL_CNOT, 0, po.FEAT_FIRST_K),
]


@pytest.mark.parametrize("case", BLK_CASES, ids=lambda c: "n%d_L%d_o%d_f%d_e%d_r%d_m%d" % c)
def test_register_block_kernels(case, eng, sample):
    """5..13 qubits run on the register-block kernels (pqc_blk.cu): forward, fitness (single block and split),
    pruned and unpruned, against the oracle and against the general shared-memory kernel."""
    import dataclasses
    n, L, n_out, nf, ent, reup, fmap = case
    spec = po.Spec(n, L, n_out, nf, ent, reup, fmap)
    ps = to_product_spec(spec)
    rng = np.random.default_rng(n * 1000 + L * 100 + n_out)
    n_cand = 5 if n <= 10 else 2
    params = rng.normal(size=(n_cand, spec.n_params)).astype(np.float32)
    n_rows = sample["sample_obs"].shape[0]
    def feats(a):
        reps = (nf + 3) // 4
        return np.ascontiguousarray(np.tile(a, (1, reps))[:, :nf])
    for B in (1, 2, 19) if n <= 10 else (1, 5):
        x = feats(sample["sample_obs"][rng.integers(0, n_rows, B)])
        ref = c_oracle.qvalues(spec, params, x)
        for variant in (ps, dataclasses.replace(ps, no_prune=1), dataclasses.replace(ps, general_kernel=1)):
            q = eng.qvalues(variant, params, x).cpu().numpy()
            assert rel_err(q, ref).max() < TOL, (B, variant)
    shapes = ((5, 1), (3, 37), (2, 700), (40, 16)) if n <= 10 else ((2, 1), (3, 21))
    for P, T in shapes:
        rows = rng.integers(0, n_rows, T)
        pp = rng.normal(size=(P, spec.n_params)).astype(np.float32)
        act = (sample["sample_act"][rows] % n_out).astype(np.int64)
        cols = [feats(sample["sample_obs"][rows]), act, sample["sample_rew"][rows], feats(sample["sample_next_obs"][rows]), sample["sample_term"][rows]]
        ref = c_oracle.fitness(spec, pp, *cols)
        f = eng.population_fitness_raw(ps, pp, *cols).cpu().numpy()
        assert rel_err(f, ref).max() < TOL, (P, T)
        f2 = eng.population_fitness_raw(ps, pp, *cols).cpu().numpy()
        assert np.array_equal(f, f2)
        g = eng.population_fitness_raw(dataclasses.replace(ps, general_kernel=1), pp, *cols).cpu().numpy()
        assert rel_err(g, ref).max() < TOL, (P, T)


def test_randomised_specs_against_oracle(eng, sample):
    """Seeded fuzz over the whole spec space (every kernel family, both entanglers, re-upload, feature maps, ragged
    population / batch sizes): Q-values and fitness against the C oracle."""
    import dataclasses
    rng = np.random.default_rng(20261020)
    n_rows = sample["sample_obs"].shape[0]
    worst = 0.0
    for case in range(70):
        n = int(rng.choice([1, 2, 3, 4, 4, 5, 6, 7, 8, 8, 9, 10, 11, 12, 13, 14, 15, 16]))
        L = int(rng.integers(0, 5 if n <= 13 else 4))
        n_out = int(rng.integers(1, min(n, 4) + 1))
        nf = int(rng.integers(1, min(n, 8) + 1))          # AngleEmbedding: at most one feature per wire
        fmap = int(rng.integers(0, 2))
        spec = po.Spec(n, L, n_out, nf, int(rng.integers(0, 2)) if n >= 2 else po.SEL_CNOT, int(rng.integers(0, 2)), fmap)
        ps = to_product_spec(spec)
        if rng.random() < 0.25:
            ps = dataclasses.replace(ps, no_prune=1)
        big = n >= 12
        P = int(rng.integers(1, 4 if big else 40))
        T = int(rng.integers(1, 6 if big else 80))
        params = rng.normal(size=(P, spec.n_params)).astype(np.float32)
        rows = rng.integers(0, n_rows, T)
        def feats(a):
            return np.ascontiguousarray(np.tile(a, (1, (nf + 3) // 4))[:, :nf])
        obs, nxt = feats(sample["sample_obs"][rows]), feats(sample["sample_next_obs"][rows])
        act = (sample["sample_act"][rows] % n_out).astype(np.int64)
        cols = [obs, act, sample["sample_rew"][rows], nxt, sample["sample_term"][rows]]
        q = eng.qvalues(ps, params, obs).cpu().numpy()
        e1 = rel_err(q, c_oracle.qvalues(spec, params, obs)).max()
        f = eng.population_fitness_raw(ps, params, *cols).cpu().numpy()
        e2 = rel_err(f, c_oracle.fitness(spec, params, *cols)).max()
        assert max(e1, e2) < TOL, (case, spec, P, T, e1, e2)
        worst = max(worst, e1, e2)
    assert worst < TOL


@pytest.mark.parametrize("ent", [po.SEL_CNOT, po.CZ_RING])
@pytest.mark.parametrize("L", [1, 2, 3, 4, 5])
def test_reference_agent_unrolled_layer_counts(L, ent, eng, sample):
    """The reference agent (4 qubits, 2 actions, 4 features) has unrolled kernels for 1..4 layers (5 = runtime loop):
    Q-values and fitness against the oracle, with the last layer folded into the read-out and without."""
    import dataclasses
    spec = po.Spec(4, L, 2, 4, ent)
    ps = to_product_spec(spec)
    rng = np.random.default_rng(40 + L)
    params = rng.normal(size=(300, spec.n_params)).astype(np.float32)
    rows = rng.integers(0, sample["sample_obs"].shape[0], 333)
    cols = [sample["sample_" + k][rows] for k in ("obs", "act", "rew", "next_obs", "term")]
    ref_q = c_oracle.qvalues(spec, params[:7], cols[0][:9])
    ref_f = c_oracle.fitness(spec, params, *cols)
    for variant in (ps, dataclasses.replace(ps, no_prune=1)):
        assert rel_err(eng.qvalues(variant, params[:7], cols[0][:9]).cpu().numpy(), ref_q).max() < TOL
        f = eng.population_fitness_raw(variant, params, *cols).cpu().numpy()
        assert rel_err(f, ref_f).max() < TOL
        assert_same_ranking(f, ref_f)


def test_ragged_and_edge_shapes(eng, sample):
    spec = to_product_spec(po.Spec())
    for P, T in ((1, 1), (1, 31), (2, 33), (25, 16), (7, 129), (300, 5), (20000, 40), (3, 4000)):
        params, rows = _c2_inputs(sample, P, T, seed=P * 1000 + T)
        cols = [sample["sample_" + k][rows] for k in ("obs", "act", "rew", "next_obs", "term")]
        f = eng.population_fitness_raw(spec, params, *cols).cpu().numpy()
        ref = po.fitness(po.Spec(), params, *cols)
        assert rel_err(f, ref).max() < TOL, (P, T)
    # odd row counts in the packed forward kernel
    for B in (1, 2, 3, 127, 257):
        params, rows = _c2_inputs(sample, 2, B, seed=B)
        q = eng.qvalues(spec, params, sample["sample_obs"][rows]).cpu().numpy()
        ref = np.stack([po.qvalues_tensor(po.Spec(), p, sample["sample_obs"][rows]) for p in params])
        assert rel_err(q, ref).max() < TOL, B
    # empty population is a no-op, empty batch is refused (mse_loss of nothing is NaN in the reference)
    import oqrl_b200
    with pytest.raises(oqrl_b200._lib.OqrlError, match="empty batch"):
        eng.population_fitness_raw(spec, np.zeros((2, 24), np.float32), np.zeros((0, 4)), np.zeros(0), np.zeros(0), np.zeros((0, 4)), np.zeros(0))
    with pytest.raises(IndexError):
        eng.population_fitness_raw(spec, np.zeros((2, 24), np.float32), np.zeros((1, 4)), np.array([2]), np.zeros(1), np.zeros((1, 4)), np.zeros(1))


def test_ga_trajectory_on_gpu_matches_reference_optim(eng, ga_ref, sample):
    """BASELINE config 1: the reference GA (P=25, G=10, B=16) with the CUDA fitness kernel reproduces the
    trajectory of the reference's own optim.py (golden, oracle-backed model) -- identical ranking in every
    generation implies bit-identical populations."""
    from oqrl_b200 import GAOptimizer, VQC
    from test_host_logic import Exp
    seed = 0
    rng = np.random.default_rng(seed)
    policy = VQC(4, 2, n_layers=2, rng=rng)
    VQC(4, 2, n_layers=2, rng=rng)
    GAOptimizer(model=policy, rng=rng)
    ga = GAOptimizer(policy, rng=rng)
    sel = rng.choice(4096, 16, replace=False)               # positions inside the committed sample
    np.testing.assert_array_equal(sample["sample_idx"][sel], ga_ref["seed0_rows"])
    batch = [Exp(torch.tensor(sample["sample_obs"][i]), torch.tensor(int(sample["sample_act"][i])), torch.tensor(float(sample["sample_rew"][i])),
                 torch.tensor(sample["sample_next_obs"][i]), torch.tensor(float(sample["sample_term"][i]))) for i in sel]
    ga.optimize(None, batch)
    np.testing.assert_array_equal(np.stack([i[0].cpu().numpy() for i in ga.population]), ga_ref["seed0_final_population"])
    np.testing.assert_array_equal(policy.params.cpu().numpy(), ga_ref["seed0_final_model_params"])
    assert ga.interaction_count == 4000


def test_fma_peak_microbench(eng):
    scalar = eng.fma_peak_tflops(0, 1 << 12)
    packed = eng.fma_peak_tflops(1, 1 << 12)
    assert 5 < scalar < 200 and 5 < packed < 200


def test_probe_library_microbenchmarks(eng):
    """The probe library's other measurements run and give plausible numbers: the register-block gate body
    (complex-packed, scalar, two-circuit, with shared-memory traffic) and the re-tiling exchange of a 16-qubit state
    in distributed shared memory vs L2 / HBM (the measurement behind DESIGN.md's choice for config 4)."""
    for form in (0, 1, 2, 6, 8):
        assert 5 < eng.gate_rate_tflops(form, 2, 256) < 200
    for medium in (0, 1, 2):
        us, gbs = eng.exchange_rate(medium, 20)
        assert 0.5 < us < 1000 and gbs > 10


# ---- HBM-resident class (n >= 15): tile passes with GF(2) layout tracking --------------------------------
HBM_CASES = [po.Spec(15, 2, 2, 4), po.Spec(15, 3, 3, 4, po.SEL_CNOT, 1, po.FEAT_TILED), po.Spec(16, 2, 2, 4, po.CZ_RING, 0, po.FEAT_TILED),
             po.Spec(16, 4, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED), po.Spec(15, 1, 2, 4), po.Spec(15, 0, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED),
             po.Spec(18, 3, 2, 4, po.CZ_RING, 1, po.FEAT_TILED), po.Spec(14, 3, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED),
             po.Spec(14, 2, 3, 4, po.CZ_RING, 0, po.FEAT_FIRST_K)]


@pytest.mark.parametrize("spec", HBM_CASES, ids=lambda s: f"n{s.n_qubits}L{s.n_layers}e{s.entangler}r{s.reupload}")
def test_hbm_class_qvalues_and_state(spec, eng, sample):
    rng = np.random.default_rng(spec.n_qubits * 10 + spec.n_layers)
    params = rng.normal(size=(2, spec.n_params)).astype(np.float32)
    x = np.ascontiguousarray(sample["sample_obs"][:3, :spec.n_feat])
    ps = to_product_spec(spec)
    assert eng.kernel_class(ps) == 2
    q = eng.qvalues(ps, params, x).cpu().numpy()
    ref = c_oracle.qvalues(spec, params, x)
    assert rel_err(q, ref).max() < TOL
    q3 = eng.qvalues(ps, params[:1], x).cpu().numpy()     # odd circuit count
    assert rel_err(q3, ref[:1]).max() < TOL
    q0 = eng.qvalues(ps, params[1:], x[2:]).cpu().numpy() # a single circuit
    assert rel_err(q0, ref[1:, 2:]).max() < TOL
    sv = eng.statevector(ps, params[0], x[0]).cpu().numpy()
    sv_ref = po.simulate(spec, params[0], x[:1])[0]
    assert np.abs(sv - sv_ref).max() < 2e-6
    assert abs(np.sum(np.abs(sv.astype(np.complex128)) ** 2) - 1.0) < 1e-5


def test_hbm_class_fitness_small(eng, sample):
    spec = po.Spec(16, 2, 2, 4)                                      # BASELINE config 4's circuit, tiny population
    rng = np.random.default_rng(4)
    params = (rng.random(spec.n_params)[None] + 0.1 * rng.normal(size=(3, spec.n_params))).astype(np.float32)
    rows = rng.choice(4096, 5, replace=False)
    cols = [sample["sample_" + k][rows] for k in ("obs", "act", "rew", "next_obs", "term")]
    ref = c_oracle.fitness(spec, params, *cols)
    ps = to_product_spec(spec)
    f_raw = eng.population_fitness_raw(ps, params, *cols).cpu().numpy()
    assert rel_err(f_raw, ref).max() < TOL
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    f_idx = eng.population_fitness(ps, params, ds, rows).cpu().numpy()
    np.testing.assert_array_equal(f_idx, f_raw)


def test_c5_24_qubit_single_circuit(eng, sample):
    """BASELINE config 5: one 24-qubit circuit, state (128 MiB complex64) resident in HBM."""
    spec = po.Spec(24, 2, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED)
    rng = np.random.default_rng(24)
    params = rng.normal(size=(1, spec.n_params)).astype(np.float32)
    x = np.ascontiguousarray(sample["sample_obs"][:1])
    ps = to_product_spec(spec)
    q = eng.qvalues(ps, params, x).cpu().numpy()
    ref = c_oracle.qvalues(spec, params, x)
    assert rel_err(q, ref).max() < TOL
    sv = eng.statevector(ps, params[0], x[0])
    norm = float((sv.real.double() ** 2 + sv.imag.double() ** 2).sum())
    assert abs(norm - 1.0) < 1e-4                                      # size-independent property: unitarity
    k = torch.arange(1 << 24, device=sv.device)
    z0 = float(((sv.real.double() ** 2 + sv.imag.double() ** 2) * (1 - 2 * ((k >> 23) & 1)).double()).sum())
    assert abs(z0 - ref[0, 0, 0]) < 1e-4                               # read-out agrees with the stored state


def test_c5_benchmarked_spec_24_qubits_8_layers(eng, sample):
    """The spec bench.py's config 5 times (n = 24, L = 8, tiled features): the whole multi-sweep plan against the C
    oracle (about 13 s of one CPU thread), plus unitarity and read-out consistency of the stored state."""
    spec = po.Spec(24, 8, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED)
    rng = np.random.default_rng(248)
    params = rng.normal(size=(1, spec.n_params)).astype(np.float32)
    x = np.ascontiguousarray(sample["sample_obs"][7:8])
    ps = to_product_spec(spec)
    q = eng.qvalues(ps, params, x).cpu().numpy()
    ref = c_oracle.qvalues(spec, params, x, nthreads=1)
    assert rel_err(q, ref).max() < TOL
    sv = eng.statevector(ps, params[0], x[0])
    prob = sv.real.double() ** 2 + sv.imag.double() ** 2
    assert abs(float(prob.sum()) - 1.0) < 1e-4
    k = torch.arange(1 << 24, device=sv.device)
    for i in range(2):
        zi = float((prob * (1 - 2 * ((k >> (23 - i)) & 1)).double()).sum())
        assert abs(zi - ref[0, 0, i]) < 1e-4


@pytest.mark.parametrize("spec", [po.Spec(20, 8, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), po.Spec(20, 8, 3, 4, po.CZ_RING, 1, po.FEAT_TILED)],
                         ids=["sel", "cz_reupload"])
def test_hbm_class_20_qubits_8_layers_full_state(spec, eng, sample):
    """A deep large circuit with every amplitude compared: 2^20 complex64 against the float64 numpy simulator."""
    rng = np.random.default_rng(208 + spec.entangler)
    params = rng.normal(size=(1, spec.n_params)).astype(np.float32)
    x = np.ascontiguousarray(sample["sample_obs"][11:12])
    ps = to_product_spec(spec)
    q = eng.qvalues(ps, params, x).cpu().numpy()
    sv_ref = po.simulate(spec, params[0], x[:1])[0]
    dim = 1 << 20
    kk = np.arange(dim)
    zref = np.array([np.sum(np.abs(sv_ref) ** 2 * (1 - 2 * ((kk >> (19 - i)) & 1))) for i in range(spec.n_out)])
    assert rel_err(q[0, 0], zref).max() < TOL
    sv = eng.statevector(ps, params[0], x[0]).cpu().numpy()
    assert np.abs(sv - sv_ref).max() < 2e-6                              # amplitudes are <= 1: absolute bound


def test_fitness_scatter_single_rank(eng, sample):
    """oqrl_fitness_scatter with this process as its only peer: the kernel's peer stores land at the block offset."""
    params, rows = _c2_inputs(sample, 300, 128, seed=3)
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    for spec in (po.Spec(), po.Spec(6, 2, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED)):
        ps = to_product_spec(spec)
        p = params if spec.n_qubits == 4 else np.random.default_rng(1).normal(size=(300, spec.n_params)).astype(np.float32)
        want = eng.population_fitness(ps, p, ds, rows)
        a = torch.full((1000,), -7.0, device="cuda")
        b = torch.full((1000,), -7.0, device="cuda")
        eng.population_fitness_scatter(ps, p, ds, rows, [a.data_ptr(), b.data_ptr()], 500)
        torch.cuda.synchronize()
        for buf in (a, b):
            assert torch.equal(buf[500:800], want)
            assert bool((buf[:500] == -7.0).all()) and bool((buf[800:] == -7.0).all())
    # the library's own cross-rank barrier with this process as its only peer: it signals itself and passes
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    for epoch in (1, 2, 3):
        eng.population_fitness_scatter(to_product_spec(po.Spec()), params, ds, rows, [a.data_ptr()], 0)
        eng.peer_barrier([flags.data_ptr()], 0, epoch)
        torch.cuda.synchronize()
        assert flags.tolist() == [epoch, 0]            # arrived, no time-out
    # host-buffer entry of the sharded evaluation (oqrl_fitness_scatter_host), this process as its only peer: pinned
    # buffers are used in place (barrier + read-back fused into one kernel), pageable ones are staged
    spec = to_product_spec(po.Spec())
    want = eng.population_fitness(spec, params, ds, rows).cpu().numpy()
    full = torch.full((1000,), -7.0, device="cuda")
    for epoch, pinned in ((4, True), (5, False)):
        hp, hi = torch.tensor(params), torch.tensor(rows.astype(np.int32))
        ho = torch.full((1000,), -3.0)
        if pinned:
            hp, hi, ho = hp.pin_memory(), hi.pin_memory(), ho.pin_memory()
        eng.population_fitness_scatter_host(spec, hp, ds, hi, [full.data_ptr()], 200, [flags.data_ptr()], 0, epoch, ho, 1000)
        np.testing.assert_array_equal(ho.numpy()[200:500], want)
        assert (ho.numpy()[:200] == -7.0).all() and (ho.numpy()[500:] == -7.0).all()
        assert flags.tolist() == [epoch, 0]
    eng.population_fitness_scatter_host(spec, torch.tensor(params).pin_memory(), ds, torch.tensor(rows.astype(np.int32)).pin_memory(),
                                        [full.data_ptr()], 0, [flags.data_ptr()], 0, 6, None, 1000)       # no read-back on this rank
    assert torch.equal(full[:300], torch.tensor(want, device="cuda")) and flags.tolist() == [6, 0]


def test_peer_scatter_two_ranks(eng):
    """Fused exchange across two GPUs (oqrl_fitness_scatter + oqrl_peer_barrier), skewed ranks, 12 generations:
    the gathered vector equals a single-GPU evaluation bit for bit.  Skipped on single-GPU boxes."""
    import subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(root, "scripts", "check_peer_scatter.py")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "peer scatter check: OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_sharded_ga_two_ranks(eng):
    """ShardedDeviceGAOptimizer across two GPUs: the reference's GA trajectory bit for bit on every rank, and an
    unevenly sharded population of 4099 against the single-GPU optimizer.  Skipped on single-GPU boxes (bench.py
    --gpus N runs the same check and reports it as `sharded_ga_matches_reference`)."""
    import subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29733", os.path.join(root, "scripts", "check_sharded_ga.py")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "sharded GA check: OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_ga_topk_selection_kernel(eng):
    """k <= 8 takes the single-block top-k kernel (what the GA uses), larger k the counting kernel: same order."""
    rng = np.random.default_rng(8)
    for n in (5, 25, 1000, 4096, 32768, 100003):
        f = rng.normal(size=n).astype(np.float32)
        f[rng.integers(0, n, size=max(1, n // 50))] = f[0]                 # exact ties
        if n > 100:
            f[rng.integers(0, n, size=3)] = np.nan
            f[rng.integers(0, n, size=2)] = np.inf
        want = np.argsort(f, kind="stable")[::-1]
        d = torch.tensor(f, device="cuda")
        for k in (1, 2, 5, 8):
            if k <= n:
                np.testing.assert_array_equal(eng.ga_rank(d, k).cpu().numpy(), want[:k], err_msg=f"n={n} k={k}")
        if n <= 4096:
            np.testing.assert_array_equal(eng.ga_rank(d, min(n, 9)).cpu().numpy(), want[:min(n, 9)])
    # the fused cross-rank wait with this process as its only peer: it signals itself and ranks
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    f = torch.tensor(rng.normal(size=300).astype(np.float32), device="cuda")
    for epoch in (1, 2):
        order = eng.ga_rank(f, 5, peers=([flags.data_ptr()], 0, epoch)).cpu().numpy()
        np.testing.assert_array_equal(order, np.argsort(f.cpu().numpy(), kind="stable")[::-1][:5])
        assert flags.tolist() == [epoch, 0]


def test_device_ga_matches_reference_trajectory(eng, ga_ref, sample):
    """population resident on the device, one breeding launch per generation, and still the
    reference optim.py trajectory bit for bit (same RNG stream, same float32 + float64-noise arithmetic)."""
    from oqrl_b200 import DeviceGAOptimizer, GAOptimizer, VQC
    from test_host_logic import Exp
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        policy = VQC(4, 2, n_layers=2, rng=rng)
        VQC(4, 2, n_layers=2, rng=rng)
        GAOptimizer(model=policy, rng=rng)
        ga = DeviceGAOptimizer(policy, rng=rng)
        sel = rng.choice(4096, 16, replace=False)
        tag = f"seed{seed}_"
        np.testing.assert_array_equal(sample["sample_idx"][sel], ga_ref[tag + "rows"])
        np.testing.assert_array_equal(np.stack([i[0].cpu().numpy() for i in ga.population]), ga_ref[tag + "init_population"])
        batch = [Exp(torch.tensor(sample["sample_obs"][i]), torch.tensor(int(sample["sample_act"][i])), torch.tensor(float(sample["sample_rew"][i])),
                     torch.tensor(sample["sample_next_obs"][i]), torch.tensor(float(sample["sample_term"][i]))) for i in sel]
        ga.optimize(None, batch)
        np.testing.assert_array_equal(np.stack([i[0].cpu().numpy() for i in ga.population]), ga_ref[tag + "final_population"])
        np.testing.assert_array_equal(policy.params.cpu().numpy(), ga_ref[tag + "final_model_params"])
        np.testing.assert_array_equal(ga.best_individual[0].cpu().numpy(), ga_ref[tag + "best_individual"])
        assert ga.interaction_count == int(ga_ref[tag + "interaction_count"])
        assert rng.random() == float(ga_ref[tag + "rng_probe"])


def test_ga_run_one_call_matches_the_stepwise_loop(eng, sample):
    """oqrl_ga_run (a whole optimize() call enqueued by one library call) leaves the same bits as driving the three
    entry points from the host -- with the reference's RNG stream and with in-kernel randomness, for the reference's
    default population and for a large one."""
    from oqrl_b200 import DeviceGAOptimizer, VQC
    from oqrl_b200.dataset import IndexBatch

    class Stepwise(DeviceGAOptimizer):            # any override of a device operation keeps the host-driven loop
        def _rank(self, fit, k):
            return eng.ga_rank(fit, k)

    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    for P, T, G, exact in ((25, 16, 10, True), (25, 16, 10, False), (700, 300, 3, True), (4099, 1024, 2, False)):
        runs = []
        for cls in (DeviceGAOptimizer, Stepwise):
            rng = np.random.default_rng(7)
            model = VQC(4, 2, n_layers=2, rng=rng)
            ga = cls(model, population_size=P, num_generations=G, rng=rng, exact_rng_stream=exact)
            batch = IndexBatch(ds, np.random.default_rng(1).choice(4096, T, replace=False))
            before = eng.launch_count()
            ga.optimize(None, batch)
            ga.optimize(None, batch)              # a second call continues from the bred population
            runs.append((ga._pop.cpu().numpy(), np.asarray(ga.last_fitness), model.params.detach().cpu().numpy(),
                         ga.interaction_count, eng.launch_count() - before, rng.random()))
        for a, b in zip(runs[0][:4], runs[1][:4]):
            np.testing.assert_array_equal(a, b)
        assert runs[0][5] == runs[1][5]           # the host generator advanced identically
        assert runs[0][4] == 2 * (3 * G + 1), runs[0][4]     # fitness + rank + breed per generation, one best-row copy per call


def test_ga_rank_kernel(eng):
    """oqrl_ga_rank = the leading entries of np.argsort(fitness)[::-1] (optim.py:207) in the order of a STABLE
    ascending sort read backwards: ties to the higher index, NaN first (np.argsort puts NaN last).  numpy's default
    sort leaves the order of exact ties unspecified (it differs between the AVX-512 and the scalar build even for a
    handful of elements), so the stable order is the one reproducible choice among the orders numpy may return."""
    rng = np.random.default_rng(5)
    for n, k in ((25, 5), (4096, 5), (1000, 1000), (70000, 8)):
        f = rng.normal(size=n).astype(np.float32)
        order = eng.ga_rank(torch.tensor(f, device="cuda"), k).cpu().numpy()
        np.testing.assert_array_equal(order, np.argsort(f, kind="stable")[::-1][:k])      # no ties in a random draw
    f = np.array([1.0, 3.0, 3.0, np.nan, -2.0, 3.0, 0.5], dtype=np.float32)
    order = eng.ga_rank(torch.tensor(f, device="cuda"), 7).cpu().numpy()
    np.testing.assert_array_equal(order, [3, 5, 2, 1, 0, 6, 4])
    # duplicates, NaN, infinities and signed zeros
    for trial in range(50):
        n = int(rng.integers(2, 17))
        f = rng.integers(-2, 3, size=n).astype(np.float32)
        f[rng.random(n) < 0.15] = np.nan
        f[rng.random(n) < 0.1] = np.inf
        f[rng.random(n) < 0.1] = -np.inf
        f[rng.random(n) < 0.1] = -0.0
        k = int(rng.integers(1, n + 1))
        order = eng.ga_rank(torch.tensor(f, device="cuda"), k).cpu().numpy()
        want = np.argsort(f, kind="stable")[::-1]
        np.testing.assert_array_equal(order, want[:k], err_msg=str(f))
        np.testing.assert_array_equal(f[order], f[np.argsort(f)[::-1][:k]])     # the reference's expression: same values rank for rank
    # large population with many duplicates: the stable order
    f = rng.integers(0, 50, size=5000).astype(np.float32)
    order = eng.ga_rank(torch.tensor(f, device="cuda"), 5000).cpu().numpy()
    np.testing.assert_array_equal(order, np.argsort(f, kind="stable")[::-1])


def test_index_and_action_range_errors(eng, sample):
    """every kernel family treats a row index outside the dataset the same way -- IndexError when the
    indices are still on the host (the reference's fancy indexing raises, dataset.py:60-66), NaN fitness when they
    are already on the device (no out-of-bounds read) -- and actions that address no read-out are rejected."""
    from oqrl_b200._lib import OqrlError
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    rng = np.random.default_rng(12)
    for spec in (po.Spec(), po.Spec(3, 2, 2, 3), po.Spec(6, 2, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED),
                 po.Spec(12, 2, 2, 4, po.SEL_CNOT, 0, po.FEAT_TILED), po.Spec(14, 2, 2, 4)):
        ps = to_product_spec(spec)
        dsx = ds if spec.n_feat == 4 else eng.DeviceDataset(sample["sample_obs"][:, :3], sample["sample_next_obs"][:, :3], sample["sample_act"],
                                                           sample["sample_rew"], sample["sample_term"])
        params = rng.normal(size=(3, spec.n_params)).astype(np.float32)
        good = np.array([5, 17, 4095, 0, 33], dtype=np.int32)
        want = eng.population_fitness(ps, params, dsx, good).cpu().numpy()
        assert np.isfinite(want).all()
        for bad_row in (-1, 4096, 2 ** 31 - 1):
            bad = good.copy()
            bad[2] = bad_row
            with pytest.raises(IndexError):
                eng.population_fitness(ps, params, dsx, bad)
            with pytest.raises(IndexError):
                eng.population_fitness_scatter(ps, params, dsx, bad, [torch.zeros(8, device="cuda").data_ptr()], 0)
            f = eng.population_fitness(ps, params, dsx, torch.tensor(bad, device="cuda")).cpu().numpy()
            assert np.isnan(f).all(), (spec, bad_row, f)
        again = eng.population_fitness(ps, params, dsx, torch.tensor(good, device="cuda")).cpu().numpy()
        np.testing.assert_array_equal(again, want)                  # the context survived: no sticky fault
    one_out = to_product_spec(po.Spec(4, 2, 1, 4))                   # the dataset holds actions 0 and 1
    with pytest.raises(OqrlError, match="actions span"):
        eng.population_fitness(one_out, rng.normal(size=(2, 24)).astype(np.float32), ds, np.arange(4, dtype=np.int32))
    wide = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"] * 3 - 1, sample["sample_rew"], sample["sample_term"])
    for spec in (po.Spec(), po.Spec(14, 2, 2, 4)):
        with pytest.raises(OqrlError, match="actions span"):
            eng.population_fitness(to_product_spec(spec), rng.normal(size=(2, spec.n_params)).astype(np.float32), wide, np.arange(4, dtype=np.int32))
    hp, ho = torch.zeros(2, 24).pin_memory(), torch.zeros(2).pin_memory()
    with pytest.raises(OqrlError, match="outside dataset"):
        eng.population_fitness_host(to_product_spec(po.Spec()), hp, ds, torch.tensor([0, 4096], dtype=torch.int32), ho)


def test_device_ga_small_population_raises(eng, sample):
    """a population smaller than the parent pool (optim.py:220 draws ranks from range(5)) is an
    IndexError up front, never an out-of-bounds read in the breeding kernel."""
    from oqrl_b200 import DeviceGAOptimizer, VQC
    from oqrl_b200._lib import OqrlError
    from oqrl_b200.dataset import IndexBatch
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    rng = np.random.default_rng(0)
    ga = DeviceGAOptimizer(VQC(4, 2, rng=rng), population_size=3, num_generations=2, rng=rng)
    with pytest.raises(IndexError):
        ga.optimize(None, IndexBatch(ds, np.arange(8, dtype=np.int32)))
    pop = torch.zeros(6, 24, device="cuda")
    with pytest.raises(OqrlError, match="d_order holds"):
        eng.ga_breed(pop, np.arange(3, dtype=np.int32), 2, np.zeros((2, 2), np.int32), np.zeros((2, 24), np.uint8), np.zeros((2, 2, 24)), parent_pool=5)


def test_ga_breed_random_distributions(eng):
    """In-kernel randomness of the vectorised GA mode: parents, mask and mutation follow optim.py:175-197."""
    P, D, elite, pool, rate, sigma = 4097, 24, 2, 5, 0.7, 0.1
    ids = np.arange(P, dtype=np.float32)
    pop = torch.tensor(np.repeat(8.0 * ids[:, None], D, axis=1), device="cuda")        # gene value identifies its parent
    order = torch.tensor(np.random.default_rng(1).permutation(P)[:pool].astype(np.int32), device="cuda")
    nxt = eng.ga_breed_random(pop, order, elite, pool, rate, sigma, seed=1234, generation=3)
    again = eng.ga_breed_random(pop, order, elite, pool, rate, sigma, seed=1234, generation=3)
    other = eng.ga_breed_random(pop, order, elite, pool, rate, sigma, seed=1234, generation=4)
    assert torch.equal(nxt, again) and not torch.equal(nxt, other)
    out, o = nxt.cpu().numpy().astype(np.float64), order.cpu().numpy()
    np.testing.assert_array_equal(out[:elite], 8.0 * o[:elite, None] * np.ones((1, D)))
    kids = out[elite:]
    parent = np.rint(kids / 8.0).astype(np.int64)
    assert np.isin(parent, o).all()
    noise = kids - 8.0 * parent
    assert abs(noise.mean()) < 2e-3 and abs(noise.std() - sigma) < 2e-3
    assert abs((np.abs(noise) > 2 * sigma).mean() - 0.0455) < 0.005          # Gaussian tails
    n_full_pairs = (P - elite) // 2                      # rows: child 1 of pair 0, child 2 of pair 0, child 1 of pair 1, ...
    c1, c2 = parent[0:2 * n_full_pairs:2], parent[1:2 * n_full_pairs:2]
    assert (c1 != c2).all()                              # complementary children of two DISTINCT parents
    lo, hi = np.minimum(c1, c2), np.maximum(c1, c2)
    assert (lo == lo[:, :1]).all() and (hi == hi[:, :1]).all()      # exactly two parents per pair
    p = (c1 == lo).mean(axis=1)                          # child 1 takes the first parent with probability `rate`
    assert abs(np.abs(p - 0.5).mean() - abs(rate - 0.5)) < 0.02
    for v in o:                                          # every pool member parents 2 / pool of the pairs
        assert abs(((lo[:, 0] == v) | (hi[:, 0] == v)).mean() - 2.0 / pool) < 0.03


def test_device_ga_vectorised_mode_improves(eng, sample):
    """exact_rng_stream=False: whole optimize() on the device (rank + in-kernel randomness); elitism keeps the best."""
    from oqrl_b200 import DeviceGAOptimizer, VQC
    from oqrl_b200.dataset import IndexBatch
    rng = np.random.default_rng(3)
    policy = VQC(4, 2, n_layers=2, rng=rng)
    ga = DeviceGAOptimizer(policy, population_size=512, num_generations=6, rng=rng, exact_rng_stream=False)
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    batch = IndexBatch(ds, np.arange(256, dtype=np.int32))
    first = max(ga.evaluate_population(batch))
    ga.optimize(None, batch)
    last = ga.last_fitness
    assert len(last) == 512 and max(last) >= first                              # elites are carried over unchanged
    assert ga.interaction_count == 512 * 256 * 7
    np.testing.assert_array_equal(policy.params.detach().cpu().numpy().reshape(-1), ga.best_individual[0].cpu().numpy().reshape(-1))


def test_es_optimizer_on_gpu(eng, sample):
    """Second metaheuristic behind BaseOptimizer on the CUDA fitness kernels."""
    from oqrl_b200 import ESOptimizer, VQC
    from oqrl_b200.dataset import IndexBatch
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    batch = IndexBatch(ds, np.arange(512, dtype=np.int32))
    rng = np.random.default_rng(9)
    policy = VQC(4, 2, n_layers=2, rng=rng)
    es = ESOptimizer(policy, mu=16, lam=240, num_generations=8, rng=rng)
    ref = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "es_trajectory_ref.npz")))
    np.testing.assert_array_equal(es.parents.cpu().numpy(), ref["init_parents"])
    first = float(eng.population_fitness(policy.spec, es.parents, ds, batch.indices).max())
    es.optimize(None, batch)
    assert es.last_fitness[0] >= first and es.interaction_count == 8 * 256 * 512
    mine = float(eng.population_fitness(policy.spec, policy.params.detach().reshape(1, -1), ds, batch.indices)[0])
    assert mine == es.last_fitness[0]                      # the model holds the best individual
    # the same run with the float64 ORACLE as the fitness backend (tests/golden/make_golden.py make_es_trajectory):
    # every selection (oqrl_ga_rank) agreed, so the surviving parents are the same bits and their fitness is within TOL
    np.testing.assert_array_equal(es.parents.cpu().numpy(), ref["final_parents"])
    assert rel_err(np.array(es.last_fitness), ref["last_fitness"]).max() < TOL
    assert es.sigma == float(ref["sigma"]) and rng.random() == float(ref["rng_probe"])


def test_ga_breed_kernel_matches_numpy(eng):
    rng = np.random.default_rng(7)
    for P, D in ((25, 24), (4096, 24), (7, 192)):
        pop = rng.normal(size=(P, D)).astype(np.float32)
        order = rng.permutation(P).astype(np.int32)
        n_pairs = -(-(P - 2) // 2)
        pair = np.stack([rng.choice(5, 2, replace=False) for _ in range(n_pairs)]).astype(np.int32)
        mask = (rng.random((n_pairs, D)) < 0.5).astype(np.uint8)
        noise = rng.normal(0, 0.1, size=(n_pairs, 2, D))
        want = [pop[order[0]], pop[order[1]]]
        for j in range(n_pairs):
            a, b = pop[order[pair[j, 0]]], pop[order[pair[j, 1]]]
            m = mask[j].astype(bool)
            want.append((np.where(m, a, b).astype(np.float64) + noise[j, 0]).astype(np.float32))
            want.append((np.where(m, b, a).astype(np.float64) + noise[j, 1]).astype(np.float32))
        want = np.stack(want)[:P]
        got = eng.ga_breed(torch.tensor(pop, device="cuda"), order[:5], 2, pair, mask, noise).cpu().numpy()
        np.testing.assert_array_equal(got, want)


def test_dataset_file_to_hbm_to_fitness(eng, sample):
    """Row a7/a8 end to end on the GPU box: HDF5 file -> OfflineDataset -> one-time HBM upload -> index batch -> fitness."""
    from oqrl_b200 import GAOptimizer, OfflineDataset, VQC
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cartpole_mini_v2layout.hdf5")
    ds = OfflineDataset(path, np.random.default_rng(11))
    dev_ds = ds.to_device()
    assert ds.to_device() is dev_ds and len(dev_ds) == 2048            # uploaded once
    idx = ds.sample_indices(64)
    model = VQC(4, 2, rng=np.random.default_rng(0))
    ga = GAOptimizer(model, population_size=25, num_generations=1, rng=np.random.default_rng(0))
    fit = ga.evaluate_population(ds.index_batch(idx))
    params = np.stack([i[0].cpu().numpy() for i in ga.population])
    ref = po.fitness(po.Spec(), params, ds.observations[idx], ds.actions[idx], ds.rewards[idx], ds.next_observations[idx], ds.terminals[idx])
    assert rel_err(np.array(fit), ref).max() < TOL
    assert ga.interaction_count == 25 * 64


def test_full_size_dataset_residency(eng, sample):
    """Row a7 at the real table size: 100 000 transitions (the committed rows tiled; the HDF5 file is not on the GPU
    box) uploaded once, then fitness on rows from the far end equals the explicit-batch path bit for bit."""
    n_rows = 100000
    reps = -(-n_rows // 4096)
    cols = {k: np.ascontiguousarray(np.concatenate([sample["sample_" + k]] * reps)[:n_rows]) for k in ("obs", "next_obs", "act", "rew", "term")}
    ds = eng.DeviceDataset(cols["obs"], cols["next_obs"], cols["act"], cols["rew"], cols["term"])
    assert len(ds) == n_rows
    rng = np.random.default_rng(100)
    rows = rng.choice(n_rows, 512, replace=False)
    rows[:4] = [0, n_rows - 1, n_rows - 4096, 65536]
    params = (rng.random(24)[None] + 0.1 * rng.normal(size=(128, 24))).astype(np.float32)
    spec = to_product_spec(po.Spec())
    f_idx = eng.population_fitness(spec, params, ds, rows.astype(np.int32)).cpu().numpy()
    f_raw = eng.population_fitness_raw(spec, params, cols["obs"][rows], cols["act"][rows], cols["rew"][rows], cols["next_obs"][rows], cols["term"][rows]).cpu().numpy()
    np.testing.assert_array_equal(f_idx, f_raw)
    ref = c_oracle.fitness(po.Spec(), params[:8], cols["obs"][rows], cols["act"][rows], cols["rew"][rows], cols["next_obs"][rows], cols["term"][rows])
    assert rel_err(f_idx[:8], ref).max() < TOL
    with pytest.raises(IndexError):
        eng.population_fitness(spec, params, ds, np.array([n_rows], dtype=np.int32))


def test_c3_full_size_properties(eng, sample):
    """BASELINE config 3 at full size: 8 qubits, 8 layers, data re-uploading, P = 16384 candidates x 1024 transitions."""
    spec = po.Spec(8, 8, 2, 4, po.SEL_CNOT, 1, po.FEAT_TILED)
    ps = to_product_spec(spec)
    rng = np.random.default_rng(3)
    P, T = 16384, 1024
    base = rng.random(spec.n_params).astype(np.float32)
    params = (base[None] + 0.1 * rng.normal(size=(P, spec.n_params))).astype(np.float32)
    rows = rng.choice(4096, T, replace=False)
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    f = eng.population_fitness(ps, params, ds, rows).cpu().numpy()
    assert np.isfinite(f).all() and (f <= 0).all()
    pick = rng.choice(P, 256, replace=False)                                  # oracle on a sample of the candidates
    ref = c_oracle.fitness(spec, params[pick], sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                           sample["sample_next_obs"][rows], sample["sample_term"][rows])
    assert rel_err(f[pick], ref).max() < TOL
    swaps = assert_same_ranking(f[pick], ref)                                 # ranking identical on the sampled sub-population
    print(f"config 3: 256 of {P} candidates ranked against the oracle, {swaps} tolerated tie swap(s), "
          f"max rel err {rel_err(f[pick], ref).max():.2e}")
    pick = pick[:6]
    # candidates are independent: a sub-population gives the same numbers bit for bit (also across grid shapes)
    sub = eng.population_fitness(ps, params[pick], ds, rows).cpu().numpy()
    assert rel_err(sub, f[pick]).max() < 1e-6


def test_c4_full_size_properties(eng, sample):
    """BASELINE config 4 at full size: 16 qubits (HBM-resident 512 KiB states), P = 1024 candidates x 256 transitions."""
    spec = po.Spec(16, 2, 2, 4)
    ps = to_product_spec(spec)
    rng = np.random.default_rng(4)
    P, T = 1024, 256
    base = rng.random(spec.n_params).astype(np.float32)
    params = (base[None] + 0.1 * rng.normal(size=(P, spec.n_params))).astype(np.float32)
    rows = rng.choice(4096, T, replace=False)
    ds = eng.DeviceDataset(sample["sample_obs"], sample["sample_next_obs"], sample["sample_act"], sample["sample_rew"], sample["sample_term"])
    f = eng.population_fitness(ps, params, ds, rows).cpu().numpy()
    assert np.isfinite(f).all() and (f <= 0).all()
    pick = rng.choice(P, 16, replace=False)
    ref = c_oracle.fitness(spec, params[pick], sample["sample_obs"][rows], sample["sample_act"][rows], sample["sample_rew"][rows],
                           sample["sample_next_obs"][rows], sample["sample_term"][rows])
    assert rel_err(f[pick], ref).max() < TOL
    swaps = assert_same_ranking(f[pick], ref)
    print(f"config 4: 16 of {P} candidates ranked against the oracle, {swaps} tolerated tie swap(s), "
          f"max rel err {rel_err(f[pick], ref).max():.2e}")


@pytest.mark.parametrize("opt", ["list", "device"])
def test_training_run_matches_reference_train_py(eng, opt):
    """Scope rows f.2/f.3: the whole offline training epoch + greedy evaluation, against the golden produced by the
    reference's OWN train.py/agent.py/optim.py/dataset.py (unmodified, run behind the pennylane/h5py/gymnasium shims)."""
    from oqrl_b200 import DeviceGAOptimizer, GAOptimizer
    from oqrl_b200.driver import run_train
    ref = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_run_ref.npz")))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cartpole_train_v2layout.hdf5")
    for seed in (0, 3):
        st = run_train(path, 1, seed, optimizer_cls=GAOptimizer if opt == "list" else DeviceGAOptimizer)
        tag = f"seed{seed}_"
        assert st["interaction_count"] == int(ref[tag + "interaction_count"]) == 8000
        assert st["update_counter"] == int(ref[tag + "update_counter"]) and st["memory_len"] == int(ref[tag + "memory_len"])
        np.testing.assert_array_equal(st["policy_params"], ref[tag + "policy_params"])           # trained weights, bit for bit
        np.testing.assert_array_equal(np.stack([i[0].cpu().numpy() for i in st["agent"].optimizer.population]), ref[tag + "population"])
        np.testing.assert_array_equal(np.array(st["eval_rewards"][0]), ref[tag + "eval_rewards"])  # 25 greedy episodes
        assert st["agent"].rng.random() == float(ref[tag + "rng_probe"])
