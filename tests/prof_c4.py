"""Developer aid (run under ncu via gpurun): a few C4 calls (1024 self-PM voices) for profiling."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads
from oracle import pyref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prg = pyref.Program(workloads.synth_c4(1024, 60))
g = saugns_b200.Generator(prg, 96000, max_call_len=24576)
for _ in range(n):
    g.run_device(24576)
print("done", g.counters())
