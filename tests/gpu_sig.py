"""Developer aid (GPU box): print the lowered-plan signature (render_fast.cuh:rec_code) of a script's first voice."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads

L = saugns_b200.lib()
for name, prg in [("c3 pm", workloads.build_c3(4096, 10, fm=False)), ("c3 fm", workloads.build_c3(4096, 10, fm=True)),
                  ("c5[0]", workloads.build_c5_script(0))]:
    g = saugns_b200.Generator(prg, 96000, max_call_len=24576)
    g.run_device(24576)
    out = (C.c_uint32 * 36)()
    L.saugen_debug_signature(C.c_void_p(g.ptr), out)
    print(name, out[0], [hex(x) for x in out[1:1 + out[0]]], flush=True)
    g.close()
