"""Developer aid (GPU box): what the first calls of a process cost (wall clock)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads
nv = int(sys.argv[1])
prg = workloads.build_c3(nv, 60, seed=1, fm=True)
t0 = time.perf_counter()
g = saugns_b200.Generator(prg, 96000, max_call_len=24576)
t1 = time.perf_counter()
ts = []
for k in range(4):
    a = time.perf_counter(); g.run_device(24576); ts.append(time.perf_counter() - a)
print(f"voices {nv}: create {1e3 * (t1 - t0):.1f} ms, calls " + " ".join(f"{1e3 * t:.2f}" for t in ts) + " ms",
      {k: os.environ.get(k) for k in ("SAUGEN_MULTI", "SAUGEN_TEAM", "CUDA_MODULE_LOADING")}, flush=True)
g.close()
