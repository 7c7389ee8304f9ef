"""Stand-in for the six PennyLane names the reference touches (agent.py:33-46,54), backed by the float64 oracle.
TEST INFRASTRUCTURE: used only by tests/golden/make_golden.py to run the reference's own files unmodified."""
import numpy as np
import torch

from oracle import pqc_oracle as po

_tape = None


class _Dev:
    def __init__(self, name, wires, shots=None):
        assert name == "default.qubit" and shots is None
        self.wires = wires


def device(name, wires, shots=None):
    return _Dev(name, wires, shots)


def AngleEmbedding(features, wires, rotation="X"):
    assert rotation == "Y"
    _tape["x"] = np.asarray(features.detach().cpu().numpy() if isinstance(features, torch.Tensor) else features, dtype=np.float32)
    _tape["n"] = len(list(wires))


def StronglyEntanglingLayers(weights, wires, ranges=None, imprimitive=None):
    assert ranges is None and imprimitive is None
    w = weights.detach().cpu().numpy() if isinstance(weights, torch.Tensor) else np.asarray(weights)
    assert w.ndim == 3 and w.shape[1] == len(list(wires)) and w.shape[2] == 3
    _tape["w"] = np.asarray(w, dtype=np.float32)


class PauliZ:
    def __init__(self, wire):
        self.wire = wire


def expval(op):
    assert isinstance(op, PauliZ)
    return op.wire


class QNode:
    def __init__(self, func, dev, interface="auto"):
        self.func, self.dev = func, dev

    def __call__(self, *args):
        global _tape
        _tape = {}
        wires = self.func(*args)
        n, w, x = _tape["n"], _tape["w"], _tape["x"]
        assert list(wires) == list(range(len(wires)))
        spec = po.Spec(n_qubits=n, n_layers=w.shape[0], n_out=len(wires), n_feat=x.shape[0])
        q = po.qvalues_tensor(spec, w.reshape(-1), x[None])[0]
        return torch.tensor(q, dtype=torch.float64)      # default.qubit + torch interface: float64 expectation values
