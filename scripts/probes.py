import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_vectordb_b200 as h
h.init(0)
pk = h.imad_peak()
print(f"imad.wide.u32 peak  {pk/1e12:.2f} T wide-MAC/s")
for w, nm, macs in ((0, "fq mul ilp1", 136), (1, "fq mul ilp2", 136), (2, "xyzz mixed add", 1360)):
    r = h.op_rate(w)
    print(f"{nm:18s} {r/1e9:8.2f} G/s  = {r*macs/1e12:.2f} T wide-MAC/s = {r*macs/pk*100:.1f}% of imad.wide peak")
