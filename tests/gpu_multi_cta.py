"""Run on the GPU box under a given SAUGEN_MULTI / SAUGEN_TEAM (tests/test_gpu_parity.py::
test_teams_over_several_ctas): few-voice scripts with nested FM, streamed at several call sizes, bit-exact."""
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

tabs = gpuutil.ref_tables_for_gpu(pyport)
feats = scripts.feature_scripts()
deep = ("Wsin f300 t1.2 p[Wtri f200.r400[Wsin f50.r90[Wsaw f7 a0.9] a0.8] a0.6] a0.5\n"
        "Wsin f220.r330[Wsin f3.r5[Wtri f0.7]] t1.1 a0.4 c0.5\n")
texts = [scripts.C2_MISC1_4FM_PM.replace("t15", "t1.5"), deep, scripts.synth_c3(40, 1.0, fm="mix"),
         feats["pm_chain"], feats["fm_both"], feats["seq_update"], feats["handover_twice"]]
for text in texts:
    prg = pyref.Program(text)
    want = pyref.render(prg, srate=96000)
    for call in (24576, 12288, 98304):
        got = saugns_b200.render(prg, srate=96000, tables=tabs, call_len=call)
        assert got.shape == want.shape and np.array_equal(got, want), (text[:40], call)
print("OK", os.environ.get("SAUGEN_MULTI"), os.environ.get("SAUGEN_TEAM"))
