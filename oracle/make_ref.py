"""Recipe for oracle/_ref/: the reference's OWN files for the hot path, copied UNMODIFIED from the read-only
reference checkout (TEST / BASELINE INFRASTRUCTURE, never product code).

The reference (frederikbckl/OQRL) is pure Python, so there is nothing to compile: "building" oracle/_ref/ means
copying the four files the path runs through -- optim.py (GAOptimizer._evaluate_fitness, optim.py:64-173),
agent.py (VQC.forward, agent.py:48-61), utils.py (Experience) and config.py -- next to a manifest of their SHA-256.
oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so the
copies travel to the GPU box, where `bench.py --impl reference` runs them behind the import shims of
tests/golden/shims (pennylane -> the float64 oracle; PennyLane itself is not installable offline).

Run by __graft_entry__.build() whenever /root/reference is present; a no-op elsewhere.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
FILES = ("optim.py", "agent.py", "utils.py", "config.py")


def build() -> bool:
    """Copy the reference's hot-path files into oracle/_ref/.  Returns False when the reference is absent."""
    if not os.path.isdir(REF_SRC):
        return False
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(REF_SRC, name)
        shutil.copyfile(src, os.path.join(OUT, name))
        manifest[name] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "sha256": manifest}, fh, indent=1)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, name)) for name in FILES)


if __name__ == "__main__":
    print("oracle/_ref:", "built" if build() else "reference checkout not present")
