"""Developer aid (GPU box): render-kernel time per call over the course of a C3 render."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads
fm = sys.argv[1] if len(sys.argv) > 1 else "mix"
fm = {"mix": "mix", "pm": False, "fm": True}[fm]
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prg = workloads.build_c3(nv, 60, seed=1, fm=fm)
g = saugns_b200.Generator(prg, 96000, max_call_len=24576)
g.set_timing(True)
prev = (0.0, 0.0)
out = []
for k in range(1, 41):
    g.run_device(24576)
    rk, mk = g.kernel_ms()
    out.append(f"{k}:{rk - prev[0]:.3f}")
    prev = (rk, mk)
print(" ".join(out), flush=True)
g.close()
