"""Developer aid (run via gpurun): wall time of BASELINE configs 1 and 2 (few voices), GPU vs reference."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import saugns_b200
import scripts
from oracle import pyref, pyport
t = pyport.ref_tables()
tabs = saugns_b200.WaveTables.from_buffer_copy(bytes(t)); tabs._keep = t
for name, text in [("C1 Wsin", "Wsin"), ("C2 misc1-4fm_pm", scripts.C2_MISC1_4FM_PM)]:
    prg = pyref.Program(text)
    saugns_b200.render(prg, srate=96000, tables=tabs, max_frames=30000)
    for call in (24576, 96000 * 4):
        t0 = time.perf_counter(); g = saugns_b200.render(prg, srate=96000, tables=tabs, call_len=call); tg = time.perf_counter() - t0
        print(f"{name}: GPU {tg:.3f} s at {call}-frame calls ({g.shape[0] / 96000 / tg:.1f}x realtime)")
    t0 = time.perf_counter(); r = pyref.render(prg, srate=96000); tr = time.perf_counter() - t0
    print(f"{name}: reference (1 core) {tr:.3f} s ({r.shape[0] / 96000 / tr:.1f}x realtime)")
