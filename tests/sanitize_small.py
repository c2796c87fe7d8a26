"""Developer aid (run under compute-sanitizer via gpurun): a small mixed workload that
touches the general interpreter, the steady-stretch plan (W / N / R / amplitude
modulators), the balanced scheduler, the mix kernel's paths and the batch calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import saugns_b200
import scripts
from saugns_b200 import batch
from oracle import pyref, pyport
t = pyport.ref_tables()
tabs = saugns_b200.WaveTables.from_buffer_copy(bytes(t)); tabs._keep = t
feats = scripts.feature_scripts()
names = ["pm_chain", "fm_both", "am_range", "noise_am", "R_cub", "self_w_mod", "pan_mod", "seq_overlap",
         "voices3", "amp_op_range", "sweep_f_cos", "Rmode_uhp", "regoal", "mod_finite"]
bad = []
for n in names:
    prg = pyref.Program(feats[n])
    want = pyref.render(prg, srate=96000)
    for call in (24576, 5000):
        got = saugns_b200.render(prg, srate=96000, tables=tabs, call_len=call)
        if got.shape != want.shape or not np.array_equal(got, want):
            bad.append((n, call))
prgs = [pyref.Program(scripts.synth_c5_script(i)) for i in range(6)]
out = batch.render_batch(prgs, srate=96000, tables=tabs, group_size=3, call_len=49152)
for p, o in zip(prgs, out):
    if not np.array_equal(o, pyref.render(p, srate=96000)):
        bad.append("batch")
prg = pyref.Program(scripts.synth_c3(96, 0.3, fm="mix"))
want = pyref.render(prg, srate=96000)
for sched in (1, 2):
    if not np.array_equal(saugns_b200.render(prg, srate=96000, tables=tabs, sched=sched), want):
        bad.append(("c3", sched))
# teams in phases (nested FM: loads / saves through the per-voice cache, counting windows), kept plans and
# the run-ahead call chain (prologue kernel, copy stream), streamed at two call sizes
c2_short = scripts.C2_MISC1_4FM_PM.replace("t15", "t0.6")
deep = ("Wsin f300 t0.7 p[Wtri f200.r400[Wsin f50.r90[Wsaw f7 a0.9] a0.8] a0.6] a0.5\n"
        "Wsin f220.r330[Wsin f3.r5[Wtri f0.7]] t0.7 a0.4 c0.5\n")
for text in (c2_short, deep):
    prg = pyref.Program(text)
    want = pyref.render(prg, srate=96000)
    for call in (24576, 8192):
        got = saugns_b200.render(prg, srate=96000, tables=tabs, call_len=call)
        if got.shape != want.shape or not np.array_equal(got, want):
            bad.append(("teams", call))
print("sanitize workload done; mismatches:", bad)
