import sys; sys.path.insert(0,'.')
from oqrl_b200 import engine
names={0:"scalar 1-fresh",1:"packed 1-fresh",2:"scalar 2-fresh",3:"packed 2-fresh",4:"scalar 3-fresh",5:"packed 3-fresh",6:"scalar 2-fresh+uniform",7:"packed 2-fresh+uniform",8:"scalar FMUL (1 flop)",9:"packed FMUL2 (1 flop)",10:"scalar FADD",11:"packed FADD2",12:"scalar 2-fresh, 4 multipliers",13:"packed 2-fresh + scalar-broadcast multiplier (Rn.F32)",14:"scalar out-of-place (dest != sources)",15:"packed out-of-place (dest != sources)"}
for v in range(16):
    r=[engine.fma_peak_tflops(v, 1<<13) for _ in range(3)]
    print(v, names[v], ["%.1f"%x for x in r])

print("packed 2-fresh at reduced occupancy (warps per SM sub-partition):")
for w in (1, 2, 3, 4, 5, 6, 8):
    print(w, "%.1f" % engine.fma_peak_tflops(3 | (w << 8), 1 << 13), "scalar:", "%.1f" % engine.fma_peak_tflops(2 | (w << 8), 1 << 13))

print("17 packed out-of-place MULTIPLY (1 flop/elem, 2 instr/round)", ["%.1f" % engine.fma_peak_tflops(17, 1 << 12) for _ in range(2)])
print("19 packed FMA with destination = multiplicand (2 instr/round)", ["%.1f" % engine.fma_peak_tflops(19, 1 << 12) for _ in range(2)])
print("packed out-of-place at reduced occupancy:")
for w in (2, 4, 8):
    print(w, "%.1f" % engine.fma_peak_tflops(15 | (w << 8), 1 << 12))
