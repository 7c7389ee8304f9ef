"""The C-ABI library loads and exports every symbol include/oqrl_b200.h declares (CPU only;
no compute entry point is called without a GPU)."""
import ctypes
import os
import re

import pytest

from oqrl_b200 import CircuitSpec, _build, _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols(header="oqrl_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oqrl_[a-z_0-9]+)\s*\(", text)))


def test_builds_in_tree():
    path = _build.build()
    assert os.path.exists(path) and os.path.dirname(path).endswith("oqrl_b200")


def test_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 16 and sorted(_lib.EXPORTS) == names
    for n in names:
        assert getattr(L, n) is not None, n


def test_probe_library_is_separate():
    """Measurement hooks (microbenchmarks, work models, pass-list dump) live in their own library: the product
    library exports none of them, the probe library exports exactly what its header declares."""
    L, P = _lib.lib(), _lib.probe()
    names = declared_symbols("oqrl_b200_probe.h")
    assert sorted(_lib.PROBE_EXPORTS) == names
    for n in names:
        assert getattr(P, n) is not None, n
        assert not hasattr(L, n), f"{n} must not be exported by the product library"
    f, s = ctypes.c_double(), ctypes.c_double()
    cs = CircuitSpec(24, 8, 2, 4, 0, 0, 1).to_c()
    assert P.oqrl_work_model(ctypes.byref(cs), ctypes.byref(f), ctypes.byref(s)) == 0 and s.value == 18 and f.value > 0


def test_spec_struct_layout():
    assert ctypes.sizeof(_lib.CSpec) == 32
    cs = CircuitSpec(8, 8, 2, 4, 0, 1, 1).to_c()
    assert (cs.n_qubits, cs.n_layers, cs.n_out, cs.n_feat, cs.entangler, cs.reupload, cs.feature_map, cs.reserved) == (8, 8, 2, 4, 0, 1, 1, 0)


def test_spec_validation_and_error_string():
    L = _lib.lib()
    assert L.oqrl_abi_version() == 1
    for bad, frag in ((CircuitSpec(0, 2, 2, 4), "n_qubits"), (CircuitSpec(4, 2, 5, 4), "n_out"),
                      (CircuitSpec(4, 2, 2, 5), "AngleEmbedding"), (CircuitSpec(4, 2, 2, 4, entangler=7), "entangler"),
                      (CircuitSpec(4, -1, 2, 4), "n_layers")):
        cs = bad.to_c()
        assert L.oqrl_kernel_class(ctypes.byref(cs)) == -1
        assert frag in L.oqrl_last_error().decode()
    for n, cls in ((1, 0), (4, 0), (5, 1), (8, 1), (13, 1), (14, 2), (16, 2), (24, 2)):
        cs = CircuitSpec(n, 2, 1, min(n, 4)).to_c()
        assert L.oqrl_kernel_class(ctypes.byref(cs)) == cls
    with pytest.raises(_lib.OqrlError, match="n_qubits"):
        _lib.check(L.oqrl_kernel_class(ctypes.byref(CircuitSpec(99, 1, 1, 1).to_c())))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oqrl_b200 import VQC, engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VQC(4, 2)(torch.zeros(1, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.qvalues(CircuitSpec(), torch.zeros(1, 24), torch.zeros(1, 4))
