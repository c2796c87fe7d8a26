"""oracle/pyport.py -- TEST INFRASTRUCTURE ONLY.

ctypes access to oracle/_ref/liboracle.so, the scalar CPU restatement of the
generator back end (oracle/saugen_oracle.cpp).  It consumes the same
`const sauProgram*` the reference front end builds (oracle/pyref.Program) or
any blob laid out per include/sau_program_abi.h.
"""
import ctypes as C
import os

import numpy as np

from . import pyref

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_ref", "liboracle.so")


class WaveTables(C.Structure):
    """Layout of saugen_WaveTables (include/saugen_b200.h)."""
    _fields_ = [("pilut", C.c_void_p * 12), ("amp_scale", C.c_float * 12),
                ("amp_dc", C.c_float * 12), ("phase_adj", C.c_int32 * 12)]


_lib = None
_ref_tables = None


def available():
    return os.path.exists(PORT_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PORT_SO, mode=os.RTLD_LOCAL)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(WaveTables)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_run.restype = C.c_int
        L.oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                 C.POINTER(C.c_size_t)]
        L.oracle_op_state.restype = C.c_int
        L.oracle_op_state.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(pyref.RefOpState)]
        L.oracle_voice_state.restype = C.c_int
        L.oracle_voice_state.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_gen_buf.restype = C.POINTER(C.c_float)
        L.oracle_gen_buf.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_mix_buf.restype = C.POINTER(C.c_float)
        L.oracle_mix_buf.argtypes = [C.c_void_p, C.c_uint32]
        _lib = L
    return _lib


def tables_from_arrays(piluts, coeffs):
    """Build a WaveTables struct (keeps the arrays alive on the struct)."""
    t = WaveTables()
    keep = np.ascontiguousarray(piluts, dtype=np.float32)
    for w in range(12):
        t.pilut[w] = keep[w].ctypes.data
        t.amp_scale[w], t.amp_dc[w], t.phase_adj[w] = coeffs[w]
    t._keep = keep
    return t


def ref_tables():
    """The reference's own host-built tables (sau/wave.c), via oracle/_ref."""
    global _ref_tables
    if _ref_tables is None:
        _ref_tables = tables_from_arrays(pyref.piluts(), pyref.picoeffs())
    return _ref_tables


class PortGenerator:
    def __init__(self, prg, srate=96000, tables=None):
        self.prg = prg
        self.tables = tables if tables is not None else ref_tables()
        self.ptr = lib().oracle_create(prg.ptr, srate, C.byref(self.tables))

    def run(self, buf_len, stereo=True):
        ch = 2 if stereo else 1
        buf = np.zeros(buf_len * ch, dtype=np.int16)
        n = C.c_size_t(0)
        more = lib().oracle_run(self.ptr, buf.ctypes.data, buf_len, int(stereo), C.byref(n))
        return bool(more), buf, n.value

    def op_state(self, op_id):
        st = pyref.RefOpState()
        lib().oracle_op_state(self.ptr, op_id, C.byref(st))
        return st

    def voice_state(self, vo_id):
        out = (C.c_uint32 * 4)()
        lib().oracle_voice_state(self.ptr, vo_id, out)
        return list(out)

    def gen_buf(self, k, n=1024):
        return np.ctypeslib.as_array(lib().oracle_gen_buf(self.ptr, k), shape=(n,)).copy()

    def close(self):
        if self.ptr:
            lib().oracle_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(prg, srate=96000, stereo=True, call_len=None, max_frames=None, tables=None):
    if call_len is None:
        call_len = srate * 256 // 1000
    g = PortGenerator(prg, srate, tables)
    chunks = []
    total = 0
    more = True
    while more:
        more, buf, n = g.run(call_len, stereo)
        chunks.append(buf[:n * (2 if stereo else 1)])
        total += n
        if max_frames and total >= max_frames:
            break
    g.close()
    ch = 2 if stereo else 1
    return np.concatenate(chunks).reshape(-1, ch) if chunks else np.zeros((0, ch), np.int16)


OP_FIELDS = ["inited", "type", "flags", "time", "i0", "i1", "mode", "oscflags", "prev_Is",
             "prev_s", "fb_s", "alpha", "rate2x"]
LINE_NAMES = ["amp", "amp2", "pan", "freq", "freq2", "pm_a"]
LINE_FIELDS = ["v0", "vt", "pos", "end", "type", "flags"]


def op_state_tuple(st):
    """Bit-exact comparable form of an op-state view (floats as raw bits)."""
    import struct

    def bits(x, fmt):
        return struct.unpack("<Q" if fmt == "d" else "<I", struct.pack("<" + fmt, x))[0]
    out = []
    for f in OP_FIELDS:
        v = getattr(st, f)
        if f == "prev_Is":
            v = bits(v, "d")
        elif f in ("prev_s", "fb_s"):
            v = bits(v, "f")
        out.append(v)
    if st.inited:
        for ln in LINE_NAMES:
            l = getattr(st, ln)
            out += [bits(l.v0, "f"), bits(l.vt, "f"), l.pos, l.end, l.type, l.flags]
    return tuple(out)
