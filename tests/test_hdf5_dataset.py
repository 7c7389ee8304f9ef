"""HDF5-subset reader and the OfflineDataset mirror (CPU only)."""
import os

import numpy as np
import pytest

from oqrl_b200 import OfflineDataset, hdf5_subset

REF_FILE = "/root/reference/offline_cartpole_test_v5.hdf5"
needs_ref = pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference data file only exists in the build container")


def test_open_failure_convention(tmp_path):
    with pytest.raises(RuntimeError, match="Failed to load dataset"):     # dataset.py:10-13
        OfflineDataset(str(tmp_path / "missing.hdf5"))
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(b"not an hdf5 file" * 64)
    with pytest.raises(RuntimeError, match="Failed to load dataset"):
        OfflineDataset(str(bad))
    with pytest.raises(hdf5_subset.Hdf5FormatError):
        hdf5_subset.File(str(bad))


@needs_ref
def test_reader_matches_appendix_c(sample):
    f = hdf5_subset.File(REF_FILE)
    assert sorted(f.keys()) == ["actions", "next_observations", "observations", "rewards", "terminals"]
    shapes = {k: (f[k].shape, f[k].dtype) for k in f.keys()}
    assert shapes["observations"] == ((100000, 4), np.dtype("<f4"))
    assert shapes["actions"] == ((100000, 1), np.dtype("<i8"))
    assert shapes["rewards"] == ((100000,), np.dtype("<f8"))
    assert shapes["terminals"] == ((100000,), np.dtype(bool))
    obs = np.array(f["observations"])
    np.testing.assert_allclose(obs.sum(axis=0, dtype=np.float64), sample["sum_obs_cols"], atol=1e-9)
    assert int(np.array(f["actions"]).sum()) == 50672
    assert int(np.array(f["terminals"]).sum()) == 2342
    idx = sample["sample_idx"]
    np.testing.assert_array_equal(obs[idx], sample["sample_obs"])
    np.testing.assert_array_equal(np.array(f["next_observations"])[:64], sample["head_next_obs"])


@needs_ref
def test_offline_dataset_mirror(sample):
    ds = OfflineDataset(REF_FILE, np.random.default_rng(0))
    assert ds.size == 100000 and ds.observations.shape == (100000, 4)
    assert ds.actions.shape == (100000, 1)                 # the reference leaves it unsqueezed on this branch
    rows = ds.sample(4096)                                 # same draw as default_rng(0).choice(N, 4096, replace=False)
    assert len(rows) == 4096 and len(rows[0]) == 5
    np.testing.assert_array_equal(np.stack([r[0] for r in rows]), sample["sample_obs"])
    np.testing.assert_array_equal(np.stack([r[3] for r in rows]), sample["sample_next_obs"])
    first = next(iter(ds))
    np.testing.assert_array_equal(first[0], sample["head_obs"][0])
    ds2 = OfflineDataset(REF_FILE, np.random.default_rng(0), squeeze_actions=True)
    assert ds2.actions.shape == (100000,)
    np.testing.assert_array_equal(ds2.sample_indices(4096), sample["sample_idx"])


MINI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cartpole_mini_v2layout.hdf5")


def test_offline_dataset_on_committed_fixture(sample):
    """The (N,1,4)/(N,1,1) layout of the missing offline_cartpole_v2.hdf5 (dataset.py:27-41 squeeze branch)."""
    f = hdf5_subset.File(MINI)
    assert f["observations"].shape == (2048, 1, 4) and f["actions"].shape == (2048, 1, 1)
    ds = OfflineDataset(MINI, np.random.default_rng(5))
    assert ds.size == 2048
    assert ds.observations.shape == (2048, 4) and ds.actions.shape == (2048,) and ds.next_observations.shape == (2048, 4)
    assert ds.rewards.dtype == np.float64 and ds.terminals.dtype == bool
    np.testing.assert_array_equal(ds.observations[:64], sample["head_obs"])
    np.testing.assert_array_equal(ds.actions[:64], sample["head_act"])
    np.testing.assert_array_equal(ds.terminals[:64], sample["head_term"])
    rows = ds.sample(16)
    idx = np.random.default_rng(5).choice(2048, size=16, replace=False)        # dataset.py:58
    np.testing.assert_array_equal(np.stack([r[0] for r in rows]), ds.observations[idx])
    assert [int(r[1]) for r in rows] == [int(a) for a in ds.actions[idx]]
    assert sum(1 for _ in ds) == 2048
