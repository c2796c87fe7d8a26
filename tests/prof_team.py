"""Developer aid (run on the GPU box): time per call of few-voice renders (team mode, render_team.cuh)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import saugns_b200
from saugns_b200 import workloads
import scripts


def time_calls(prg, n=20, frames=24576):
    g = saugns_b200.Generator(prg, 96000, max_call_len=frames)
    for _ in range(6):             # past the scripts' first second (C3: a slower call where its 1 s ramps end)
        g.run_device(frames)
    g.set_timing(True)
    t0 = time.perf_counter()
    for _ in range(n):
        g.run_device(frames)
    wall = (time.perf_counter() - t0) / n
    rk, mk = g.kernel_ms()
    g.close()
    return rk / n, mk / n, wall * 1e3


for fm in (False, True, "mix"):
    for nv in (1, 64, 512, 1024, 2048, 4096):
        prg = workloads.build_c3(nv, 60, seed=1, fm=fm)
        rk, mk, wall = time_calls(prg)
        print(f"C3 fm={fm!s:5} voices={nv:5d} render {rk:7.3f} ms  mix {mk:6.3f} ms  call {wall:7.3f} ms", flush=True)
try:
    from oracle import pyref
    prg = pyref.Program(scripts.C2_MISC1_4FM_PM)
    t0 = time.perf_counter()
    pcm = saugns_b200.render(prg, srate=96000)
    t1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = pyref.render(prg, srate=96000)
    t2 = time.perf_counter() - t0
    print(f"C2 misc1-4fm_pm: GPU {t1:.3f} s, reference (one core) {t2:.3f} s, equal {np.array_equal(pcm, ref)}")
    rk, mk, wall = time_calls(prg, 40)
    print(f"C2 per call: render {rk:.3f} ms mix {mk:.3f} ms call {wall:.3f} ms")
except Exception as e:
    print("C2 skipped:", e)
