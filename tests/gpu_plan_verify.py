"""Run on the GPU box with SAUGEN_PLAN_VERIFY=1 (tests/test_gpu_parity.py::test_kept_plans_equal_fresh_ones):
every time a voice's kept plan would have been used, it is compared with the plan built afresh."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import saugns_b200
import scripts
from oracle import pyref, pyport
import gpuutil

assert os.environ.get("SAUGEN_PLAN_VERIFY") == "1"
tabs = gpuutil.ref_tables_for_gpu(pyport)
L = saugns_b200.lib()
feats = scripts.feature_scripts()
texts = [scripts.synth_c3(64, 1.5, fm="mix"), scripts.C2_MISC1_4FM_PM, scripts.synth_c5_script(3), scripts.synth_c5_script(7),
         feats["seq_update"], feats["pm_chain"], feats["fm_both"], feats["handover_twice"], feats["regoal"],
         feats["am_range_sweep"], feats["wave_change"], feats["voices3"]]
total = [0, 0]
for text in texts:
    prg = pyref.Program(text)
    want = pyref.render(prg, srate=96000)
    for call in (24576, 4096):
        g = saugns_b200.Generator(prg, 96000, tables=tabs, max_call_len=call)
        chunks, more = [], True
        while more:
            more, buf, n = g.run(call)
            chunks.append(buf[:2 * n].copy())
        got = np.concatenate(chunks).reshape(-1, 2)
        assert np.array_equal(got, want), text[:40]
        out = (C.c_uint32 * 32)()
        L.saugen_debug_team(C.c_void_p(g.ptr), out)
        g.close()
        total = [out[30], out[31]]
print(f"kept plans compared {total[1]}, differing {total[0]}")
assert total[1] > 100 and total[0] == 0
print("OK")
