"""Developer aid (run under ncu via gpurun): a few calls of many voices of ONE line of script."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from oracle import pyref, pyport

t = pyport.ref_tables()
tabs = saugns_b200.WaveTables.from_buffer_copy(bytes(t))
tabs._keep = t
line = sys.argv[1]
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rnd = random.Random(3)
text = f"S a.m{0.3 / nv ** 0.5:.6f}\n" + "".join(
    line.replace("{f}", f"{110.0 * 2 ** rnd.uniform(0, 4):.3f}").replace("{c}", f"{rnd.uniform(-1, 1):.3f}") + "\n"
    for _ in range(nv))
g = saugns_b200.Generator(pyref.Program(text), 96000, tables=tabs, max_call_len=24576)
for _ in range(4):
    g.run_device(24576)
print("done", g.counters())
