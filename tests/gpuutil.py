"""Helpers shared by the GPU parity tests."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Golden:
    """Committed answers of the unmodified reference (tests/golden/*.json), applied
    only where the wave tables of the host under test are the ones the answers were
    made with (tests/golden/make_golden.py, "HOST DEPENDENCE")."""

    def __init__(self, pyref):
        import glob
        t = pyref.piluts()
        mine = {w: hashlib.sha256(t[i].tobytes()).hexdigest() for i, w in enumerate(pyref.WAVES)}
        self.variants = []
        for path in sorted(glob.glob(os.path.join(HERE, "golden", "known_answers*.json"))):
            with open(path) as f:
                d = json.load(f)
            tabs = d.get("_meta", {}).get("tables", {})
            ok_mask = 0
            for i, w in enumerate(pyref.WAVES):
                if tabs.get(w) == mine[w]:
                    ok_mask |= 1 << i
            self.variants.append((ok_mask, d))
        self.variants.sort(key=lambda v: -bin(v[0]).count("1"))
        self.applied = 0

    def get(self, key):
        """The answer for `key` from a variant whose tables match for every wave the
        script uses, else None (the live-reference comparison still covers it)."""
        for ok_mask, d in self.variants:
            e = d.get(key)
            if e is not None and (e.get("waves", 0xfff) & ~ok_mask) == 0:
                self.applied += 1
                return e
        return None

    def __getitem__(self, key):
        e = self.get(key)
        if e is None:
            import pytest
            pytest.skip(f"no golden answer for {key} made with this host's wave tables")
        return e


def golden(pyref=None):
    if pyref is None:
        from oracle import pyref
    return Golden(pyref)


def ref_tables_for_gpu(pyport):
    """The reference's own host-built tables, as a saugen_WaveTables."""
    from saugns_b200.generator import WaveTables
    t = pyport.ref_tables()
    w = WaveTables.from_buffer_copy(bytes(t))
    w._keep = t
    return w


def sha(pcm):
    return hashlib.sha256(np.ascontiguousarray(pcm).tobytes()).hexdigest()
