"""Developer aid (run under ncu via gpurun): a few C3 steps for profiling."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
voices = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prg = workloads.build_c3(voices, 60, seed=1, fm="mix")
g = saugns_b200.Generator(prg, 96000, max_call_len=24576)
for _ in range(n):
    g.run_device(24576)
print("done", g.counters())
