"""Developer aid (run under gpurun): GPU vs reference on feature scripts with
a per-script mismatch report.  Not collected by pytest."""
import sys
import os
import ctypes as C
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import scripts
import saugns_b200
from saugns_b200.generator import WaveTables
from oracle import pyref, pyport


def ref_tables():
    t = pyport.ref_tables()
    w = WaveTables.from_buffer_copy(bytes(t))
    w._keep = t
    return w


def main():
    sel = sys.argv[1] if len(sys.argv) > 1 else ""
    feats = scripts.feature_scripts()
    tabs = ref_tables()
    bad = 0
    for name, text in feats.items():
        if sel and sel not in name:
            continue
        prg = pyref.Program(text)
        want = pyref.render(prg, srate=96000)
        try:
            got = saugns_b200.render(prg, srate=96000, tables=tabs)
        except Exception as e:
            print("EXC", name, e)
            bad += 1
            continue
        if got.shape != want.shape:
            print(f"SHAPE {name}: got {got.shape} want {want.shape}")
            bad += 1
            continue
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        nd = int((d != 0).sum())
        if nd:
            bad += 1
            first = int(np.nonzero(d.max(axis=1))[0][0])
            print(f"DIFF {name}: {nd} differ, max {d.max()}, first at {first}/{want.shape[0]} [{text}]")
    print("feature scripts:", len(feats), "bad:", bad)


if __name__ == "__main__":
    main()
