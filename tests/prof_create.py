import os, sys, time
sys.path.insert(0, "/root/repo")
import saugns_b200
from saugns_b200 import workloads
prg = workloads.build_c3(4096, 60, seed=1, fm="mix")
for i in range(4):
    t0 = time.perf_counter()
    g = saugns_b200.Generator(prg, 96000, max_call_len=24576)
    t1 = time.perf_counter()
    g.run(24576)
    t2 = time.perf_counter()
    g.close()
    t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms, first call {1e3*(t2-t1):.2f} ms, destroy {1e3*(t3-t2):.2f} ms", flush=True)
