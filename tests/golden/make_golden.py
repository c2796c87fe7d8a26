"""Generates tests/golden/known_answers.json from the UNMODIFIED reference
(oracle/_ref/libsauref.so, built from /root/reference by oracle/Makefile).

For every script: frame count and sha256 of the 16-bit PCM the reference
renders at the stated rate/channels, plus the integer oscillator state
(phase / cycle / counter words) of every operator at the end.  Run from the
repo root in the container that has /root/reference:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import scripts  # noqa: E402
from oracle import pyref  # noqa: E402


def entry(text, srate=96000, stereo=True):
    prg = pyref.Program(text)
    g = pyref.RefGenerator(prg, srate)
    call = srate * 256 // 1000
    h = hashlib.sha256()
    frames = 0
    more = True
    ch = 2 if stereo else 1
    while more:
        more, buf, n = g.run(call, stereo)
        h.update(buf[:n * ch].tobytes())
        frames += n
    state = []
    for op in range(prg.op_count):
        st = g.op_state(op)
        state.append([st.inited, st.type, st.i0, st.i1, st.time])
    return {"srate": srate, "stereo": stereo, "frames": frames, "sha256": h.hexdigest(),
            "op_state": state, "vo_count": prg.vo_count, "op_count": prg.op_count}


def main():
    out = {}
    for name, text in scripts.feature_scripts().items():
        out["feat/" + name] = entry(text)
    out["config/C1_Wsin"] = entry("Wsin")
    out["config/C2_misc1_4fm_pm"] = entry(scripts.C2_MISC1_4FM_PM)
    out["config/C3_64v_1s"] = entry(scripts.synth_c3(64, 1))
    out["config/C3fm_64v_1s"] = entry(scripts.synth_c3(64, 1, fm=True))
    out["config/C4_48v_1s"] = entry(scripts.synth_c4(48, 1))
    for i in range(8):
        out[f"config/C5_script{i}"] = entry(scripts.synth_c5_script(i))
    out["mono/voices3"] = entry(scripts.feature_scripts()["voices3"], 44100, False)
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
