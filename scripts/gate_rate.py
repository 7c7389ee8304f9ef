"""Probe library: register-block gate body rates (TFLOP/s at 28 flop per pair and gate) -- see oqrl_gate_rate."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oqrl_b200 import engine
forms = {0: "registers only, uniform coef", 1: "scalar FFMA", 2: "two-circuit packing", 3: "+ smem ld/st", 6: "+ smem, XOR addresses",
         8: "+ smem, XOR addresses, vector coef", 9: "+ smem, XOR addresses, RZ-RY-RZ split with 16-entry phase tables (same gates)"}
sel = [int(a) for a in sys.argv[1:]] or sorted(forms)
for f in sel:
    print(f, forms.get(f, ""), [(w, round(engine.gate_rate_tflops(f, w, 4096), 1)) for w in (1, 2, 3)])
