import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sample():
    """Committed rows of offline_cartpole_test_v5.hdf5 (tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "cartpole_v5_sample.npz")))


@pytest.fixture(scope="session")
def golden_fitness():
    return dict(np.load(os.path.join(GOLDEN, "golden_fitness.npz")))


@pytest.fixture(scope="session")
def ga_ref():
    return dict(np.load(os.path.join(GOLDEN, "ga_trajectory_ref.npz")))


@pytest.fixture(scope="session")
def kat():
    import json
    return json.load(open(os.path.join(GOLDEN, "kat_appendix_b.json")))


def to_product_spec(ospec):
    """oracle Spec -> oqrl_b200.CircuitSpec (same fields)."""
    from oqrl_b200 import CircuitSpec
    return CircuitSpec(ospec.n_qubits, ospec.n_layers, ospec.n_out, ospec.n_feat,
                       ospec.entangler, ospec.reupload, ospec.feature_map)


def rel_err(a, ref):
    """|a-ref| / max(|ref|, 1)  -- the SURVEY.md 8(d) correctness gate (Q crosses zero)."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)


def assert_same_ranking(fit_gpu, fit_ref, tie_rel=2e-6):
    """np.argsort(f)[::-1] identical (optim.py:207), up to swaps inside groups of
    candidates whose ORACLE fitness differs by less than `tie_rel` (relative):
    fp32 arithmetic cannot order those and numpy's default argsort is unstable.
    Tie policy documented in DESIGN.md.  Returns the number of rank positions that differ (all inside such
    tie groups), so that callers can report it instead of passing silently."""
    fit_gpu, fit_ref = np.asarray(fit_gpu, np.float64), np.asarray(fit_ref, np.float64)
    og, orf = np.argsort(fit_gpu)[::-1], np.argsort(fit_ref)[::-1]
    if np.array_equal(og, orf):
        return 0
    bad = np.nonzero(og != orf)[0]
    for pos in bad:
        a, b = fit_ref[og[pos]], fit_ref[orf[pos]]
        assert abs(a - b) <= tie_rel * max(abs(a), abs(b), 1.0), \
            f"rank {pos}: gpu picks cand {og[pos]} (oracle {a}), oracle picks {orf[pos]} ({b})"
    return int(len(bad))
