"""Developer aid (run via gpurun): render-kernel time per voice kind of the C5 mix
(many voices of ONE kind per script), to find the slow paths of the interpreter."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads
from saugns_b200.workloads import WAVES, NOISES, LINES
from oracle import pyref, pyport

t = pyport.ref_tables()
tabs = saugns_b200.WaveTables.from_buffer_copy(bytes(t))
tabs._keep = t
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
CALL = 98304


def script(kind_fn, seed=5):
    rnd = random.Random(seed)
    lines = [f"S a.m{0.3 / NV ** 0.5:.6f}"]
    for _ in range(NV):
        lines.append(kind_fn(rnd, 110.0 * 2 ** rnd.uniform(0, 4), rnd.uniform(-1, 1)))
    return "\n".join(lines) + "\n"


kinds = {}
kinds["W pm (2 ops)"] = lambda r, f, c: (f"W{r.choice(WAVES)} f{f:.3f} t4 c{c:.3f} p[W{r.choice(WAVES)} "
                                         f"r{r.choice([0.5, 1, 2, 3])} a{r.uniform(0.1, 1):.3f}]")
kinds["W pm sin only"] = lambda r, f, c: f"Wsin f{f:.3f} t4 c{c:.3f} p[Wsin r2 a{r.uniform(0.1, 1):.3f}]"
for nz in NOISES:
    kinds[f"N{nz}"] = (lambda nz: lambda r, f, c: f"N{nz} t4 c{c:.3f} a{r.uniform(0.1, 0.8):.3f}")(nz)
for m in "ugbtfa":
    kinds[f"Rlin m{m}"] = (lambda m: lambda r, f, c: f"Rlin m{m} f{f:.3f} t4 c{c:.3f}")(m)
for fl in "hpsvz":
    kinds[f"Rlin mu{fl}"] = (lambda fl: lambda r, f, c: f"Rlin mu{fl} f{f:.3f} t4 c{c:.3f}")(fl)
for ln in LINES:
    kinds[f"R{ln} mu"] = (lambda ln: lambda r, f, c: f"R{ln} mu f{f:.3f} t4 c{c:.3f}")(ln)
for ln in LINES:
    kinds[f"W sweep l{ln} + range-AM"] = (lambda ln: lambda r, f, c: (
        f"Wsin f{f:.3f}[g{f * r.uniform(0.5, 2):.3f} l{ln}] t4 c{c:.3f} a1.r0[Wsin f{r.uniform(0.5, 9):.3f}]"))(ln)

for name, fn in kinds.items():
    prg = pyref.Program(script(fn))
    g = saugns_b200.Generator(prg, 96000, tables=tabs, max_call_len=CALL)
    g.run_device(CALL)
    g.set_timing(True)
    for _ in range(2):
        g.run_device(CALL)
    rk, mk = g.kernel_ms()
    g.close()
    vs = NV * CALL * 2
    print(f"{name:28s} render {rk / 2:8.3f} ms/call  {vs / (rk * 1e-3) / 1e9:7.2f} G voice-samples/s")
