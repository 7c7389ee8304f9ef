"""ctypes binding of include/oqrl_b200.h.  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os

from . import _build

_lib = None

# every symbol include/oqrl_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "oqrl_abi_version", "oqrl_last_error", "oqrl_device_info",
    "oqrl_dataset_upload", "oqrl_dataset_free", "oqrl_dataset_shape", "oqrl_dataset_gather",
    "oqrl_qvalues", "oqrl_fitness", "oqrl_fitness_raw", "oqrl_fitness_host", "oqrl_fitness_scatter", "oqrl_fitness_scatter_host", "oqrl_peer_barrier", "oqrl_ga_breed", "oqrl_ga_rank", "oqrl_ga_rank_peers", "oqrl_ga_breed_random", "oqrl_ga_run",
    "oqrl_statevector", "oqrl_kernel_class", "oqrl_kernel_name", "oqrl_launch_count",
]
# every symbol include/oqrl_b200_probe.h declares (measurement / test hooks, a separate library)
PROBE_EXPORTS = ["oqrl_work_model", "oqrl_hbm_plan", "oqrl_reg_plan", "oqrl_fma_peak", "oqrl_gate_rate", "oqrl_exchange_rate"]
_probe = None


class OqrlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oqrl_b200 error {code}: {msg}")
        self.code = code


class CSpec(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in
                ("n_qubits", "n_layers", "n_out", "n_feat", "entangler", "reupload", "feature_map", "reserved")]


def lib() -> ctypes.CDLL:
    """Load liboqrl_b200.so (building it in-tree with nvcc if absent)."""
    global _lib
    if _lib is None:
        path = os.environ.get("OQRL_B200_LIB") or _build.LIB_PATH   # override = A/B builds of the same sources
        if not os.path.exists(path):
            path = _build.build()
        L = ctypes.CDLL(path)
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
        sp = ctypes.POINTER(CSpec)
        L.oqrl_abi_version.restype = ctypes.c_int
        L.oqrl_last_error.restype = ctypes.c_char_p
        L.oqrl_kernel_name.restype = ctypes.c_char_p
        L.oqrl_launch_count.restype = i64
        L.oqrl_device_info.argtypes = [ctypes.POINTER(i32)] * 3
        L.oqrl_dataset_upload.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp, ctypes.POINTER(vp)]
        L.oqrl_dataset_free.argtypes = [vp]
        L.oqrl_dataset_shape.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i32)]
        L.oqrl_dataset_gather.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
        L.oqrl_qvalues.argtypes = [sp, vp, i32, vp, i32, vp, vp]
        L.oqrl_fitness.argtypes = [sp, vp, i32, vp, vp, i32, f32, vp, vp]
        L.oqrl_fitness_raw.argtypes = [sp, vp, i32, vp, vp, vp, vp, vp, i32, f32, vp, vp]
        L.oqrl_fitness_host.argtypes = [sp, vp, i32, vp, vp, i32, f32, vp, vp]
        L.oqrl_ga_breed.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, vp, vp, i32, vp, vp]
        L.oqrl_ga_rank.argtypes = [vp, i32, i32, vp, vp]
        L.oqrl_ga_breed_random.argtypes = [vp, i32, i32, vp, i32, i32, f32, f32, ctypes.c_uint64, ctypes.c_uint32, vp, vp]
        L.oqrl_fitness_scatter.argtypes = [sp, vp, i32, vp, vp, i32, f32, ctypes.POINTER(vp), i32, i32, vp]
        L.oqrl_peer_barrier.argtypes = [ctypes.POINTER(vp), i32, i32, ctypes.c_uint32, vp]
        L.oqrl_fitness_scatter_host.argtypes = [sp, vp, i32, vp, vp, i32, f32, ctypes.POINTER(vp), i32, i32, ctypes.POINTER(vp), i32,
                                                ctypes.c_uint32, vp, i32, vp]
        L.oqrl_ga_rank_peers.argtypes = [vp, i32, i32, vp, ctypes.POINTER(vp), i32, i32, ctypes.c_uint32, vp]
        L.oqrl_ga_run.argtypes = [sp, vp, vp, i32, vp, vp, i32, f32, i32, i32, i32, vp, vp, vp, i32, f32, f32, ctypes.c_uint64,
                                  ctypes.c_uint32, vp, vp, vp, ctypes.POINTER(i32), vp]
        L.oqrl_statevector.argtypes = [sp, vp, vp, vp, vp]
        L.oqrl_kernel_class.argtypes = [sp]
        L.oqrl_kernel_name.argtypes = [sp]
        for name in EXPORTS:
            if name not in ("oqrl_last_error", "oqrl_launch_count", "oqrl_kernel_name"):
                getattr(L, name).restype = ctypes.c_int
        if L.oqrl_abi_version() != 1:
            raise OqrlError(-1, "ABI version mismatch")
        _lib = L
    return _lib


def probe() -> ctypes.CDLL:
    """Load liboqrl_b200_probe.so (bench.py / tests only: microbenchmarks, work models, the HBM-class pass list)."""
    global _probe
    if _probe is None:
        lib()                                                   # the probe library links against the product library
        P = ctypes.CDLL(_build.PROBE_LIB_PATH)
        vp, i32 = ctypes.c_void_p, ctypes.c_int32
        sp, dp = ctypes.POINTER(CSpec), ctypes.POINTER(ctypes.c_double)
        P.oqrl_work_model.argtypes = [sp, dp, dp]
        P.oqrl_fma_peak.argtypes = [i32, i32, dp, vp]
        P.oqrl_hbm_plan.argtypes = [sp, vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), i32]
        P.oqrl_reg_plan.argtypes = [i32, i32, i32, i32, ctypes.c_int64, ctypes.POINTER(i32)]
        P.oqrl_gate_rate.argtypes = [i32, i32, i32, dp, vp]
        P.oqrl_exchange_rate.argtypes = [i32, i32, dp, dp, vp]
        for name in PROBE_EXPORTS:
            getattr(P, name).restype = ctypes.c_int
        _probe = P
    return _probe


def check(rc: int) -> None:
    if rc != 0:
        raise OqrlError(rc, lib().oqrl_last_error().decode(errors="replace"))
