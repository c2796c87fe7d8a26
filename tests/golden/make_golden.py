"""Generates tests/golden/known_answers.json from the UNMODIFIED reference
(oracle/_ref/libsauref.so, built from /root/reference by oracle/Makefile).

For every script: frame count and sha256 of the 16-bit PCM the reference
renders at the stated rate/channels, plus the integer oscillator state
(phase / cycle / counter words) of every operator at the end.  Run from the
repo root in the container that has /root/reference:
    python tests/golden/make_golden.py [out.json]

HOST DEPENDENCE.  The reference builds its wave tables on the host with
`-O3 -ffast-math` (sau/wave.c:77-221), and gcc vectorises the libm calls there
into libmvec, whose last-bit rounding depends on the CPU's ISA dispatch: the
srs/cat/mto tables (and so every answer that uses them) differ between the
AMD EPYC container and the Xeon GPU boxes.  Each golden file therefore carries
the sha256 of the 12 tables it was made with (`_meta.tables`) and, per entry,
the waves the script uses; tests/gpuutil.golden() applies an answer only when
the tables of the host under test match.  One file per table variant:
known_answers.epyc.json (made in the build container), known_answers.xeon.json (made on a GPU
box: `python tests/golden/make_golden.py gpurun_out/known_answers.box.json`).
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


def wave_mask(prg):
    """Waves whose tables a program can read (sin is every W operator's initial wave)."""
    from saugns_b200 import program as P
    m = 0
    for ev in P.dump(prg.ptr)["events"]:
        for od in ev["ops"]:
            if od["type"] == P.POPT_WAVE:
                m |= 1
                if od["params"] & P.POPP_MODE:
                    m |= 1 << od["mode"][1]
    return m


def table_hashes():
    t = pyref.piluts()
    return {w: hashlib.sha256(t[i].tobytes()).hexdigest() for i, w in enumerate(pyref.WAVES)}


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
            "op_state": state, "vo_count": prg.vo_count, "op_count": prg.op_count,
            "waves": wave_mask(prg)}


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
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out["_meta"] = {"tables": table_hashes(), "host_cpu": cpu}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "known_answers.epyc.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out) - 1, "entries to", path)


if __name__ == "__main__":
    main()
