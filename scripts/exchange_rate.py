"""Probe library: exposed cost of re-tiling one 16-qubit state between two passes, in distributed shared memory of an
eight-CTA cluster vs L2 / HBM (oqrl_exchange_rate) -- the measurement behind DESIGN.md's choice for config 4."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oqrl_b200 import engine  # noqa: E402
names = {0: "DSMEM, cp.async.bulk smem -> peer smem", 1: "L2/HBM, cp.async.bulk store + load", 2: "DSMEM, st.shared::cluster"}
for m in (0, 1, 2):
    us, gbs = engine.exchange_rate(m, 400)
    print(f"medium {m} ({names[m]}): {us:.2f} us per 512 KiB exchange x 18 clusters, {gbs:.0f} GB/s aggregate payload")
