"""Helpers shared by the GPU parity tests."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with open(os.path.join(HERE, "golden", "known_answers.json")) as f:
        return json.load(f)


def ref_tables_for_gpu(pyport):
    """The reference's own host-built tables, as a saugen_WaveTables."""
    from saugns_b200.generator import WaveTables
    t = pyport.ref_tables()
    w = WaveTables.from_buffer_copy(bytes(t))
    w._keep = t
    return w


def sha(pcm):
    return hashlib.sha256(np.ascontiguousarray(pcm).tobytes()).hexdigest()
