#!/usr/bin/env python
"""Writes saugns_b200/data/sau_wave_tables.bin: the 12 pre-integrated wave tables and the
per-wave coefficients exactly as the reference's front-end library builds them on the host
(sau_global_init_Wave, sau/wave.c:105-221; sauWave_piluts / sauWave_picoeffs, sau/wave.c:49-66,
sau/wave.h:33-70), read out of oracle/_ref/libsauref.so.  SURVEY.md section 8a, a13: the tables
are INPUT DATA of the generator path ("upload the host-built tables verbatim"); this file is
what callers without libsau in their process (bench.py's product arm, the Python host)
pass to saugen_create.  Run by __graft_entry__.build() where /root/reference exists.

Layout (little endian): magic "SAUT", u32 version = 1, u32 waves = 12, u32 len = 2048,
12 x 2048 f32 tables, 12 f32 amp_scale, 12 f32 amp_dc, 12 i32 phase_adj."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "saugns_b200", "data", "sau_wave_tables.bin")


def main():
    from oracle import pyref
    t = pyref.piluts()
    c = pyref.picoeffs()
    blob = struct.pack("<4sIII", b"SAUT", 1, 12, 2048) + t.astype("<f4").tobytes()
    blob += np.array([x[0] for x in c], "<f4").tobytes()
    blob += np.array([x[1] for x in c], "<f4").tobytes()
    blob += np.array([x[2] for x in c], "<i4").tobytes()
    old = open(OUT, "rb").read() if os.path.exists(OUT) else None
    if old != blob:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        with open(OUT, "wb") as f:
            f.write(blob)
        print(f"wrote {OUT} ({len(blob)} bytes)" + ("" if old is None else " (content changed)"))
    else:
        print(f"{OUT} up to date")


if __name__ == "__main__":
    main()
