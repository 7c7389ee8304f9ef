"""Drop-in ``VQC`` (reference: /root/reference/src/agent.py:14-61) on the CUDA path.

Same constructor, attributes, parameter layout and seeded initialisation; the
per-sample PennyLane QNode loop (agent.py:51-58) is replaced by one
``oqrl_qvalues`` launch for the whole batch.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import engine
from .spec import CircuitSpec, ENT_SEL_CNOT, FEAT_FIRST_K


class VQC(nn.Module):
    """Variational quantum circuit Q-function: x[B, input_dim] -> <Z_i>[B, output_dim].

    Extra keyword arguments (defaults = the reference circuit) expose the
    builder-defined generalisations of SURVEY.md Appendix A.7: ``n_qubits`` (more
    wires than features), ``entangler``, ``reupload`` and ``feature_map``.
    """

    def __init__(self, input_dim, output_dim, n_layers=2, rng=None, *, n_qubits=None,
                 entangler=ENT_SEL_CNOT, reupload=False, feature_map=FEAT_FIRST_K):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.n_layers = n_layers
        self.rng = rng or np.random.default_rng()
        self.n_qubits = int(n_qubits) if n_qubits is not None else int(input_dim)
        self.spec = CircuitSpec(n_qubits=self.n_qubits, n_layers=int(n_layers), n_out=int(output_dim),
                                n_feat=int(input_dim), entangler=int(entangler), reupload=int(bool(reupload)),
                                feature_map=int(feature_map))
        # agent.py:30-31 -- uniform [0,1) draws from the shared generator, float32, no grad
        init_values = self.rng.random(self.spec.n_params).astype(np.float32)
        self.params = nn.Parameter(torch.tensor(init_values), requires_grad=False)

    def forward(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in x]) if len(x) else \
                torch.zeros((0, self.input_dim), dtype=torch.float32)
        x = x.reshape(-1, self.input_dim) if x.ndim != 2 else x
        if x.shape[0] == 0:
            raise RuntimeError("stack expects a non-empty TensorList")  # torch.stack([]) in agent.py:61
        q = engine.qvalues(self.spec, self.params.data, x)[0]
        # the QNode returns float64 expectation values; keep the dtype
        return q.to(torch.float64)
