"""Turns the reference's own script corpus (/root/reference/examples/**,
devtests/**) into a fixture that travels to the GPU box, which has no /root/reference:
tests/golden/ref_corpus.json = for every script its path, its text (INPUT DATA of the parity
tests, saugns v0.4.7 by Joel K. Pettersson, quoted verbatim like tests/scripts.py does for
BASELINE config 2), the sample rate the tests render it at, and what the unmodified reference
(oracle/_ref) makes of it on this host: frames, sha256 of the PCM, the waves it uses and the
table hashes the answer belongs to (tests/golden/make_golden.py, "HOST DEPENDENCE").
Scripts the reference's own front end cannot build are listed with "builds": false
(devtests/crashes/* segfault its parser: rc 139 from `saugns -c`; they have no output to
compare).  Run where /root/reference exists:  python tests/golden/make_corpus.py
"""
import glob
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference"
from oracle import pyref  # noqa: E402
import make_golden  # noqa: E402

LONG_MS = 150000          # longer scripts are rendered at 8 kHz (devtests/alarm-25m.sau: 25 minutes)


def main():
    out = {"_meta": {"tables": make_golden.table_hashes(), "source": "saugns v0.4.7 examples/ and devtests/"},
           "scripts": []}
    files = sorted(glob.glob(os.path.join(REF, "examples", "**", "*.sau"), recursive=True) +
                   glob.glob(os.path.join(REF, "devtests", "**", "*.sau"), recursive=True))
    for path in files:
        rel = os.path.relpath(path, REF)
        text = open(path).read()
        e = {"path": rel, "text": text}
        # the reference CLI's own verdict first: a crashing parser must not take this process down
        r = subprocess.run([pyref.REF_EXE, "-c", "-d", path], capture_output=True)
        e["check_rc"] = r.returncode
        if r.returncode != 0:
            e["builds"] = False
            out["scripts"].append(e)
            continue
        try:
            prg = pyref.Program(text)
        except ValueError:
            e["builds"] = False
            out["scripts"].append(e)
            continue
        e["builds"] = True
        srate = 96000 if prg.duration_ms <= LONG_MS else 8000
        ans = make_golden.entry(text, srate)
        e.update({"srate": srate, "frames": ans["frames"], "sha256": ans["sha256"], "waves": ans["waves"],
                  "vo_count": ans["vo_count"], "op_count": ans["op_count"], "duration_ms": prg.duration_ms})
        out["scripts"].append(e)
    with open(os.path.join(HERE, "ref_corpus.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    ok = sum(1 for e in out["scripts"] if e["builds"])
    print(f"{len(out['scripts'])} scripts, {ok} build, total {sum(e.get('frames', 0) for e in out['scripts'])} frames")


if __name__ == "__main__":
    main()
