"""ctypes loader for oracle/pqc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used by tests (cross-check of the numpy oracle) and by bench.py's CPU-baseline
legs (multi-threaded "best effort" CPU row).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboqrl_oracle.so")
_lib = None


class CSpec(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in
                ("n_qubits", "n_layers", "n_out", "n_feat", "entangler", "reupload", "feature_map", "reserved")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pqc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_qvalues.restype = ctypes.c_int
        _lib.oracle_fitness.restype = ctypes.c_int
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _cspec(spec) -> CSpec:
    return CSpec(spec.n_qubits, spec.n_layers, spec.n_out, spec.n_feat,
                 spec.entangler, spec.reupload, spec.feature_map, 0)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def qvalues(spec, params, x, nthreads=0):
    params = np.atleast_2d(_f32(params))
    x = np.atleast_2d(_f32(x))
    P, B = params.shape[0], x.shape[0]
    q = np.empty((P, B, spec.n_out), dtype=np.float64)
    cs = _cspec(spec)
    rc = lib().oracle_qvalues(ctypes.byref(cs), params.ctypes.data_as(ctypes.c_void_p), P,
                              x.ctypes.data_as(ctypes.c_void_p), B,
                              q.ctypes.data_as(ctypes.c_void_p), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_qvalues failed: {rc}")
    return q


def fitness(spec, params, obs, act, rew, next_obs, term, gamma=0.99, nthreads=0):
    params = np.atleast_2d(_f32(params))
    obs, next_obs, rew, term = _f32(obs), _f32(next_obs), _f32(rew).reshape(-1), _f32(term).reshape(-1)
    act = np.ascontiguousarray(np.asarray(act).reshape(-1), dtype=np.int32)
    P, T = params.shape[0], obs.shape[0]
    out = np.empty(P, dtype=np.float64)
    cs = _cspec(spec)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().oracle_fitness(ctypes.byref(cs), vp(params), P, vp(obs), vp(next_obs), vp(act),
                              vp(rew), vp(term), T, ctypes.c_double(gamma), vp(out), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_fitness failed: {rc}")
    return out


def max_threads() -> int:
    return int(lib().oracle_max_threads())
