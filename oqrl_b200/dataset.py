"""Drop-in ``OfflineDataset`` (reference: /root/reference/src/dataset.py) with HBM residency.

Host-side behaviour (attributes, squeeze rules, seeded ``sample``, ``__iter__``,
the ``RuntimeError("Failed to load dataset: ...")`` convention) follows the
reference; ``to_device()`` uploads the whole table to HBM once and
``sample_indices`` / ``IndexBatch`` are the index-returning fast path that
feeds ``GAOptimizer`` without materialising per-row tuples.
"""
from __future__ import annotations

import numpy as np

from . import hdf5_subset


class IndexBatch:
    """A batch that is a list of row indices into an HBM-resident dataset.

    ``len(batch)`` is the number of transitions, so it can be handed to
    ``GAOptimizer.optimize`` wherever the reference passes ``list[Experience]``.
    """

    def __init__(self, dataset, indices):
        self.dataset = dataset            # engine.DeviceDataset
        self.indices = np.ascontiguousarray(indices, dtype=np.int32) if not hasattr(indices, "is_cuda") else indices

    def __len__(self):
        return int(self.indices.shape[0])


class OfflineDataset:
    """Offline dataset of pre-collected transitions."""

    def __init__(self, file_path, rng=None, squeeze_actions=False):
        try:
            self.file = hdf5_subset.File(file_path, "r")
            obs = np.array(self.file["observations"])
            act = np.array(self.file["actions"])
            rew = np.array(self.file["rewards"])
            nxt = np.array(self.file["next_observations"])
            term = np.array(self.file["terminals"])
        except Exception as e:  # dataset.py:10-13
            raise RuntimeError(f"Failed to load dataset: {e}")
        self.rng = rng or np.random.default_rng()
        if obs.shape == (100000, 4):
            # dataset.py:22-26: arrays are used as stored; `actions` stays (N, 1) there
            # (SURVEY.md Appendix C "known reference defect"); squeeze_actions=True fixes it.
            self.observations, self.actions, self.rewards, self.next_observations = obs, act, rew, nxt
            if squeeze_actions:
                self.actions = act.reshape(len(act))
        else:
            # dataset.py:27-41: (N,1,4) observations and (N,1,1) actions lose their unit axes
            self.observations = np.squeeze(obs, axis=1)
            self.actions = np.squeeze(act, axis=(1, 2))
            self.rewards = rew
            self.next_observations = np.squeeze(nxt, axis=1)
        self.terminals = term
        self.size = len(self.observations)
        self._device = None

    def __iter__(self):
        for i in range(self.size):
            yield (self.observations[i], self.actions[i], self.rewards[i],
                   self.next_observations[i], self.terminals[i])

    def sample_indices(self, size):
        """The seeded draw of dataset.py:58 on its own (same generator state afterwards)."""
        return self.rng.choice(self.size, size=size, replace=False)

    def sample(self, size):
        """Seeded subset as a list of row tuples (dataset.py:56-66)."""
        idx = self.sample_indices(size)
        return list(zip(self.observations[idx], self.actions[idx], self.rewards[idx],
                        self.next_observations[idx], self.terminals[idx]))

    # -- HBM residency ------------------------------------------------------------------
    def to_device(self):
        """Upload the whole table to HBM once (idempotent) -> engine.DeviceDataset."""
        if self._device is None:
            from . import engine
            self._device = engine.DeviceDataset(self.observations, self.next_observations,
                                                np.asarray(self.actions).reshape(self.size),
                                                self.rewards, self.terminals)
        return self._device

    def index_batch(self, indices):
        return IndexBatch(self.to_device(), indices)
