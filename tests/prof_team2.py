"""Developer aid (GPU box): what the teams did on the last stretch (render_team.cuh:g_team_dump)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import saugns_b200
from saugns_b200 import workloads
import scripts

L = saugns_b200.lib()


def show(name, prg, calls=8, frames=24576):
    g = saugns_b200.Generator(prg, 96000, max_call_len=frames)
    for _ in range(calls):
        g.run_device(frames)
    g.set_timing(True)
    n = 10
    for _ in range(n):
        g.run_device(frames)
    rk, mk = g.kernel_ms()
    out = (C.c_uint32 * 32)()
    L.saugen_debug_team(C.c_void_p(g.ptr), out)
    o = list(out)
    print(f"{name}: render {rk / n:.3f} ms/call; offered {o[0]} split {o[16]} P {o[1]} t_eff {o[2]} nrec {o[3]} C {o[4]} "
          f"slots {o[5]} analyse {o[6]} cyc, stretch {o[7]} cyc ({o[7] / 1965e3:.3f} ms), phases {o[8:8 + 8]}\n"
          f"    time line of warp 0 (cycles since kernel start): staged {o[18]} voice {o[19]} ops+events {o[20]} plan {o[21]} "
          f"lowered {o[22]} matched {o[23]} rendered {o[24]} updated {o[25]} stored {o[26]}", flush=True)
    g.close()


from oracle import pyref
show("C2", pyref.Program(scripts.C2_MISC1_4FM_PM))
show("C3 pm x1", workloads.build_c3(1, 60, seed=1, fm=False))
show("C3 fm x1", workloads.build_c3(1, 60, seed=1, fm=True))
show("C3 mix x512", workloads.build_c3(512, 60, seed=1, fm="mix"))
show("C3 mix x2048", workloads.build_c3(2048, 60, seed=1, fm="mix"))
show("C3 mix x4096", workloads.build_c3(4096, 60, seed=1, fm="mix"))
