"""In-tree nvcc build of liboqrl_b200.so (sm_100a only; cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "liboqrl_b200.so")
PROBE_LIB_PATH = os.path.join(_HERE, "liboqrl_b200_probe.so")     # measurement / test hooks, not the product
SOURCES = ["api.cu", "pqc_reg.cu", "pqc_smem.cu", "pqc_blk.cu", "pqc_hbm.cu"]
PROBE_SOURCES = ["probe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
    "--threads", "0",          # the five translation units compile in parallel
]


def _nvcc() -> str:
    for cand in (os.environ.get("OQRL_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; oqrl_b200 has no prebuilt or CPU fallback")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(PROBE_LIB_PATH):
        return True
    t = min(os.path.getmtime(LIB_PATH), os.path.getmtime(PROBE_LIB_PATH))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(_HERE, "..", "include", "oqrl_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH + ".tmp", *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    # the probe library resolves the planner / work models from the product library next to it
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", PROBE_LIB_PATH + ".tmp", *[os.path.join(CSRC, s) for s in PROBE_SOURCES],
           "-L" + _HERE, "-l:liboqrl_b200.so", "-Xlinker", "-rpath=$ORIGIN"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    os.replace(PROBE_LIB_PATH + ".tmp", PROBE_LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
