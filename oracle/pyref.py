"""oracle/pyref.py -- TEST INFRASTRUCTURE ONLY.

ctypes access to oracle/_ref/libsauref.so: the UNMODIFIED reference (parser +
generator, built by oracle/Makefile with the reference's own flags) plus the
white-box harness oracle/ref_harness.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libsauref.so")
REF_EXE = os.path.join(_HERE, "_ref", "saugns_ref")

# reference enums (sau/program.h:69-80, sau/wave.h:33-70, sau/line.h:18-32,
# sau/program.h:102-110)
POPT = {"amp": 0, "noise": 1, "wave": 2, "raseg": 3}
WAVES = ["sin", "tri", "srs", "sqr", "ean", "cat", "eto", "par", "mto", "saw", "hsi", "spa"]
LINES = ["cos", "lin", "sah", "exp", "log", "xpe", "lge", "sqe", "cub", "smo", "ncl", "nhl", "uwh"]
NOISES = ["wh", "gw", "bw", "tw", "re", "vi", "bv"]


class RefLineState(C.Structure):
    _fields_ = [("v0", C.c_float), ("vt", C.c_float), ("pos", C.c_uint32),
                ("end", C.c_uint32), ("type", C.c_uint32), ("flags", C.c_uint32)]


class RefOpState(C.Structure):
    _fields_ = [("inited", C.c_uint32), ("type", C.c_uint32), ("flags", C.c_uint32),
                ("time", C.c_uint32),
                ("amp", RefLineState), ("amp2", RefLineState), ("pan", RefLineState),
                ("freq", RefLineState), ("freq2", RefLineState), ("pm_a", RefLineState),
                ("i0", C.c_uint32), ("i1", C.c_uint32), ("mode", C.c_uint32),
                ("oscflags", C.c_uint32), ("prev_Is", C.c_double),
                ("prev_s", C.c_float), ("fb_s", C.c_float),
                ("alpha", C.c_uint32), ("rate2x", C.c_uint32)]


_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libsauref.so missing: run `make -C oracle ref` "
                               "where /root/reference exists")
        L = C.CDLL(REF_SO, mode=os.RTLD_LOCAL)
        L.refwb_build_program.restype = C.c_void_p
        L.refwb_build_program.argtypes = [C.c_char_p, C.c_int]
        L.refwb_discard_program.argtypes = [C.c_void_p]
        L.refwb_program_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.refwb_create.restype = C.c_void_p
        L.refwb_create.argtypes = [C.c_void_p, C.c_uint32]
        L.refwb_destroy.argtypes = [C.c_void_p]
        L.refwb_run.restype = C.c_int
        L.refwb_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                C.POINTER(C.c_size_t)]
        L.refwb_render.restype = C.c_size_t
        L.refwb_render.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_size_t,
                                   C.c_void_p, C.c_size_t]
        L.refwb_render_null.restype = C.c_size_t
        L.refwb_render_null.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_size_t,
                                        C.c_size_t]
        L.refwb_op_state.restype = C.c_int
        L.refwb_op_state.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(RefOpState)]
        L.refwb_voice_state.restype = C.c_int
        L.refwb_voice_state.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.refwb_gen_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.refwb_amp_scale.restype = C.c_float
        L.refwb_amp_scale.argtypes = [C.c_void_p]
        L.refwb_gen_buf.restype = C.POINTER(C.c_float)
        L.refwb_gen_buf.argtypes = [C.c_void_p, C.c_uint32]
        L.refwb_mix_buf.restype = C.POINTER(C.c_float)
        L.refwb_mix_buf.argtypes = [C.c_void_p, C.c_uint32]
        L.refwb_pilut.restype = C.POINTER(C.c_float)
        L.refwb_pilut.argtypes = [C.c_uint32]
        L.refwb_lut.restype = C.POINTER(C.c_float)
        L.refwb_lut.argtypes = [C.c_uint32]
        L.refwb_picoeffs.argtypes = [C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_int32)]
        L.refwb_line_fill.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_float,
                                      C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        L.refwb_line_map.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                     C.c_void_p]
        L.refwb_line_val.restype = C.c_float
        L.refwb_line_val.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float]
        L.refwb_abi_layout.restype = C.c_size_t
        L.refwb_abi_layout.argtypes = [C.POINTER(C.c_uint32), C.c_size_t]
        _lib = L
    return _lib


class Program:
    """A sauProgram built by the reference front end (sau/parser.c:2093)."""

    def __init__(self, script, is_path=False):
        s = script.encode() if isinstance(script, str) else script
        self._keep = s
        self.ptr = lib().refwb_build_program(s, 1 if is_path else 0)
        if not self.ptr:
            raise ValueError("reference parser rejected the script")
        info = (C.c_uint32 * 6)()
        lib().refwb_program_info(self.ptr, info)
        (self.ev_count, self.vo_count, self.op_count, self.op_nest_depth,
         self.duration_ms, self.mode) = list(info)

    def close(self):
        if self.ptr:
            lib().refwb_discard_program(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefGenerator:
    """Streaming wrapper over the reference generator (sau/generator.h:20-26)."""

    def __init__(self, prg, srate=96000):
        self.prg = prg
        self.srate = srate
        self.ptr = lib().refwb_create(prg.ptr, srate)
        if not self.ptr:
            raise MemoryError("sau_create_Generator returned NULL")

    def run(self, buf_len, stereo=True):
        ch = 2 if stereo else 1
        buf = np.zeros(buf_len * ch, dtype=np.int16)
        n = C.c_size_t(0)
        more = lib().refwb_run(self.ptr, buf.ctypes.data, buf_len, int(stereo), C.byref(n))
        return bool(more), buf, n.value

    def op_state(self, op_id):
        st = RefOpState()
        if lib().refwb_op_state(self.ptr, op_id, C.byref(st)) != 0:
            raise IndexError(op_id)
        return st

    def voice_state(self, vo_id):
        out = (C.c_uint32 * 4)()
        lib().refwb_voice_state(self.ptr, vo_id, out)
        return list(out)

    def gen_buf(self, k, n=1024):
        return np.ctypeslib.as_array(lib().refwb_gen_buf(self.ptr, k), shape=(n,)).copy()

    def mix_buf(self, ch, n=1024):
        return np.ctypeslib.as_array(lib().refwb_mix_buf(self.ptr, ch), shape=(n,)).copy()

    @property
    def amp_scale(self):
        return lib().refwb_amp_scale(self.ptr)

    def close(self):
        if self.ptr:
            lib().refwb_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(script, srate=96000, stereo=True, call_len=None, is_path=False, max_frames=None):
    """Render a whole script with the reference; returns int16 [frames, ch]."""
    prg = script if isinstance(script, Program) else Program(script, is_path)
    if call_len is None:
        call_len = srate * 256 // 1000          # saugns.c:471,526-527
    ch = 2 if stereo else 1
    cap = (prg.duration_ms * srate + 999) // 1000 + 2 * call_len
    if max_frames:
        cap = min(cap, max_frames)
    out = np.zeros(cap * ch, dtype=np.int16)
    n = lib().refwb_render(prg.ptr, srate, int(stereo), call_len, out.ctypes.data, cap)
    return out[:n * ch].reshape(n, ch)


def piluts():
    """The 12 pre-integrated wave tables as built by the reference (sau/wave.c:49-62)."""
    t = np.zeros((len(WAVES), 2048), dtype=np.float32)
    for w in range(len(WAVES)):
        t[w] = np.ctypeslib.as_array(lib().refwb_pilut(w), shape=(2048,))
    return t


def picoeffs():
    out = []
    for w in range(len(WAVES)):
        a, d, p = C.c_float(), C.c_float(), C.c_int32()
        lib().refwb_picoeffs(w, C.byref(a), C.byref(d), C.byref(p))
        out.append((a.value, d.value, p.value))
    return out


def abi_layout():
    buf = (C.c_uint32 * 128)()
    n = lib().refwb_abi_layout(buf, 128)
    return list(buf[:n])
