"""oqrl_b200 -- B200-native (sm_100a) fitness path for OQRL's PQC Q-function agents.

Host-side mirrors of the reference interfaces (``VQC``, ``GAOptimizer``,
``BaseOptimizer``, ``OfflineDataset``) over hand-written CUDA kernels reached
through the C ABI in ``include/oqrl_b200.h``.  Importing the package does not
need a GPU; calling a kernel without one raises.
"""
from .spec import CircuitSpec, ENT_CZ_RING, ENT_SEL_CNOT, FEAT_FIRST_K, FEAT_TILED  # noqa: F401
from .vqc import VQC  # noqa: F401
from .optim import BaseOptimizer, DeviceGAOptimizer, ESOptimizer, GAOptimizer, ShardedDeviceGAOptimizer  # noqa: F401
from .dataset import IndexBatch, OfflineDataset  # noqa: F401
from . import engine  # noqa: F401

__all__ = ["CircuitSpec", "VQC", "BaseOptimizer", "GAOptimizer", "DeviceGAOptimizer", "ShardedDeviceGAOptimizer", "ESOptimizer", "OfflineDataset", "IndexBatch", "engine",
           "ENT_SEL_CNOT", "ENT_CZ_RING", "FEAT_FIRST_K", "FEAT_TILED"]
