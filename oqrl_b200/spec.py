"""Circuit description shared by the host-side mirrors (== ``oqrl_spec``)."""
from __future__ import annotations

from dataclasses import dataclass

from ._lib import CSpec

ENT_SEL_CNOT = 0   # qml.StronglyEntanglingLayers default (agent.py:45)
ENT_CZ_RING = 1
FEAT_FIRST_K = 0   # qml.AngleEmbedding (agent.py:44)
FEAT_TILED = 1


@dataclass(frozen=True)
class CircuitSpec:
    n_qubits: int = 4
    n_layers: int = 2
    n_out: int = 2
    n_feat: int = 4
    entangler: int = ENT_SEL_CNOT
    reupload: int = 0
    feature_map: int = FEAT_FIRST_K
    no_prune: int = 0   # measurement aid: 1 disables the exact last-layer light-cone pruning (oqrl_spec.reserved bit 0)
    general_kernel: int = 0   # measurement aid: 1 runs the 5..13-qubit class on its general kernel (oqrl_spec.reserved bit 1)
    direct_loads: int = 0   # measurement aid: 1 keeps the 1..4-qubit fitness kernel from staging the batch in shared memory (bit 2)

    @property
    def n_params(self) -> int:
        return self.n_layers * self.n_qubits * 3

    def to_c(self) -> CSpec:
        return CSpec(self.n_qubits, self.n_layers, self.n_out, self.n_feat,
                     self.entangler, self.reupload, self.feature_map, (int(self.no_prune) & 1) | ((int(self.general_kernel) & 1) << 1) | ((int(self.direct_loads) & 1) << 2))
