"""Developer aid (run under gpurun): ticketed scheduler on feature scripts, one by one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import scripts, gpuutil
import saugns_b200
from oracle import pyref, pyport
tabs = gpuutil.ref_tables_for_gpu(pyport)
feats = scripts.feature_scripts()
names = sys.argv[1:] or ["seq_update", "voices3", "regoal", "seq_overlap", "silence_mid", "pm_addrem"]
for name in names:
    prg = pyref.Program(feats[name])
    print(name, "voices", prg.vo_count, "ops", prg.op_count, flush=True)
    want = pyref.render(prg, srate=96000)
    got = saugns_b200.render(prg, srate=96000, tables=tabs, sched=2, call_len=8192)
    print("  equal:", got.shape == want.shape and np.array_equal(got, want), flush=True)
