"""ctypes binding of libsaugen_b200.so and the Generator class.

`Generator(prg, srate)` / `.run(buf_len, stereo)` / `.close()` mirror
sau_create_Generator / sauGenerator_run / sau_destroy_Generator
(reference sau/generator.c:200-228,905-973): same argument meaning, same
return convention (`more`, PCM, `out_len`), NULL -> exception on failure.
`prg` is anything with a `.ptr` attribute holding the address of a sauProgram
laid out per include/sau_program_abi.h (the reference front end's output, or
saugns_b200.program.ProgramBuilder).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsaugen_b200.so")


class WaveTables(C.Structure):
    """saugen_WaveTables (include/saugen_b200.h)."""
    _fields_ = [("pilut", C.c_void_p * 12), ("amp_scale", C.c_float * 12),
                ("amp_dc", C.c_float * 12), ("phase_adj", C.c_int32 * 12)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("voice_begin", C.c_uint32),
                ("voice_end", C.c_uint32), ("max_call_len", C.c_uint32), ("sched", C.c_uint32),
                ("pcm_big_endian", C.c_uint32)]


class LineView(C.Structure):
    _fields_ = [("v0", C.c_float), ("vt", C.c_float), ("pos", C.c_uint32),
                ("end", C.c_uint32), ("type", C.c_uint32), ("flags", C.c_uint32)]


class OpView(C.Structure):
    """saugen_OpView; same layout as oracle/ref_harness.c:RefOpState."""
    _fields_ = [("inited", C.c_uint32), ("type", C.c_uint32), ("flags", C.c_uint32),
                ("time", C.c_uint32),
                ("amp", LineView), ("amp2", LineView), ("pan", LineView),
                ("freq", LineView), ("freq2", LineView), ("pm_a", LineView),
                ("i0", C.c_uint32), ("i1", C.c_uint32), ("mode", C.c_uint32),
                ("oscflags", C.c_uint32), ("prev_Is", C.c_double),
                ("prev_s", C.c_float), ("fb_s", C.c_float),
                ("alpha", C.c_uint32), ("rate2x", C.c_uint32)]


_lib = None


def lib():
    """Load the CUDA extension; fail loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). saugns_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
        L.saugen_create.restype = C.c_void_p
        L.saugen_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.saugen_destroy.argtypes = [C.c_void_p]
        L.saugen_flatten.restype = C.c_size_t
        L.saugen_flatten.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        L.saugen_create_flat.restype = C.c_void_p
        L.saugen_create_flat.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.saugen_run.restype = C.c_int
        L.saugen_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                 C.POINTER(C.c_size_t)]
        L.saugen_run_device.restype = C.c_int
        L.saugen_run_device.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t)]
        L.saugen_run_many.restype = C.c_int
        L.saugen_run_many.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int,
                                      C.c_void_p, C.c_void_p]
        L.saugen_batch_create.restype = C.c_void_p
        L.saugen_batch_create.argtypes = [C.c_int]
        L.saugen_batch_destroy.argtypes = [C.c_void_p]
        L.saugen_batch_begin.restype = C.c_int
        L.saugen_batch_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.c_int, C.c_int]
        L.saugen_batch_end.restype = C.c_int
        L.saugen_batch_end.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.saugen_pinned_alloc.restype = C.c_void_p
        L.saugen_pinned_alloc.argtypes = [C.c_size_t]
        L.saugen_pinned_free.argtypes = [C.c_void_p]
        L.saugen_run_mix.restype = C.c_int
        L.saugen_run_mix.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t)]
        L.saugen_mix_to_pcm.restype = C.c_int
        L.saugen_mix_to_pcm.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.saugen_read_op.restype = C.c_int
        L.saugen_read_op.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(OpView)]
        L.saugen_read_voice.restype = C.c_int
        L.saugen_read_voice.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.saugen_read_voice_rows.restype = C.c_int
        L.saugen_read_voice_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                             C.c_size_t]
        L.saugen_counters.restype = C.c_int
        L.saugen_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.saugen_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.saugen_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.saugen_selftest.restype = C.c_longlong
        L.saugen_selftest.argtypes = [C.c_int, C.c_void_p]
        L.saugen_amp_scale.restype = C.c_float
        L.saugen_amp_scale.argtypes = [C.c_void_p]
        L.saugen_last_error.restype = C.c_char_p
        L.saugen_device_count.restype = C.c_int
        L.saugen_wave_tables_load.restype = C.POINTER(WaveTables)
        L.saugen_wave_tables_load.argtypes = [C.c_char_p]
        L.saugen_wave_tables_free.argtypes = [C.c_void_p]
        L.saugen_render_batch.restype = C.c_int
        L.saugen_render_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.saugen_render_batch_wav.restype = C.c_int
        L.saugen_render_batch_wav.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.saugen_batch_last_error.restype = C.c_char_p
        L.saugen_voice_groups.restype = C.c_int
        L.saugen_voice_groups.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.saugen_abi_layout.restype = C.c_size_t
        L.saugen_abi_layout.argtypes = [C.POINTER(C.c_uint32), C.c_size_t]
        _lib = L
    return _lib


def last_error():
    return lib().saugen_last_error().decode()


def device_count():
    return lib().saugen_device_count()


def selftest(device=0, tables=None):
    """Mismatches of the fast-path arithmetic primitives vs their plain forms (0 = exact)."""
    if tables is None:
        tables = default_tables()
    r = lib().saugen_selftest(device, C.addressof(tables))
    if r < 0:
        raise RuntimeError("saugen_selftest failed: " + last_error())
    return r


TABLES_PATH = os.path.join(_HERE, "data", "sau_wave_tables.bin")
_default_tables = None


def default_tables():
    """libsau's host-built wave tables (sau/wave.c:105-221), read verbatim from the data file
    tools/make_wave_tables.py wrote: what a caller without libsau in its process passes to
    saugen_create (the drop-in passes libsau's live arrays instead, csrc/dropin.c)."""
    global _default_tables
    if _default_tables is None:
        p = lib().saugen_wave_tables_load(TABLES_PATH.encode())
        if not p:
            raise RuntimeError("wave tables: " + last_error())
        _default_tables = p.contents        # lives for the process
    return _default_tables


def tables_arrays(t=None):
    """(piluts[12,2048] float32, [(amp_scale, amp_dc, phase_adj)]*12) of a WaveTables."""
    t = t if t is not None else default_tables()
    tabs = np.zeros((12, 2048), np.float32)
    for w in range(12):
        tabs[w] = np.ctypeslib.as_array(C.cast(t.pilut[w], C.POINTER(C.c_float)), shape=(2048,))
    return tabs, [(t.amp_scale[w], t.amp_dc[w], t.phase_adj[w]) for w in range(12)]


def abi_layout():
    buf = (C.c_uint32 * 128)()
    n = lib().saugen_abi_layout(buf, 128)
    return list(buf[:n])


class Generator:
    """One sauGenerator instance living on a B200."""

    def __init__(self, prg, srate=96000, tables=None, device=0, stream=None,
                 voice_range=None, max_call_len=0, sched=0, big_endian=False):
        self._prg = prg            # borrowed for the generator's life (generator.c:191)
        if tables is None:
            tables = default_tables()
        self._tables = tables
        flat = prg if isinstance(prg, (bytes, bytearray)) else None   # a saugen_flatten blob
        opt = Options(device=device, stream=stream or 0,
                      voice_begin=voice_range[0] if voice_range else 0,
                      voice_end=voice_range[1] if voice_range else 0,
                      max_call_len=max_call_len, sched=sched, pcm_big_endian=int(big_endian))
        tptr = C.addressof(tables)
        if flat is not None:
            buf = (C.c_char * len(flat)).from_buffer_copy(flat)
            self.ptr = lib().saugen_create_flat(buf, len(flat), tptr, C.byref(opt))
        else:
            self.ptr = lib().saugen_create(prg.ptr, srate, tptr, C.byref(opt))
        if not self.ptr:
            raise RuntimeError("saugen_create failed: " + last_error())
        self.srate = srate

    def run(self, buf_len, stereo=True):
        """-> (more, int16 array of buf_len*channels, out_len)."""
        ch = 2 if stereo else 1
        buf = np.zeros(buf_len * ch, dtype=np.int16)
        n = C.c_size_t(0)
        r = lib().saugen_run(self.ptr, buf.ctypes.data, buf_len, int(stereo), C.byref(n))
        if r < 0:
            raise RuntimeError("saugen_run failed: " + last_error())
        return bool(r), buf, n.value

    def run_device(self, buf_len, stereo=True):
        """Render one call leaving the PCM in HBM -> (more, device pointer, out_len)."""
        n = C.c_size_t(0)
        p = C.c_void_p(0)
        r = lib().saugen_run_device(self.ptr, buf_len, int(stereo), C.byref(p), C.byref(n))
        if r < 0:
            raise RuntimeError("saugen_run_device failed: " + last_error())
        return bool(r), p.value, n.value

    def run_mix(self, buf_len):
        """Partial float mix planes in HBM (voice-sharded multi-GPU) -> (more, dev ptr, out_len)."""
        n = C.c_size_t(0)
        p = C.c_void_p(0)
        r = lib().saugen_run_mix(self.ptr, buf_len, C.byref(p), C.byref(n))
        if r < 0:
            raise RuntimeError("saugen_run_mix failed: " + last_error())
        return bool(r), p.value, n.value

    def mix_to_pcm(self, dev_mix_ptr, buf_len, stereo=True):
        ch = 2 if stereo else 1
        buf = np.zeros(buf_len * ch, dtype=np.int16)
        if lib().saugen_mix_to_pcm(self.ptr, dev_mix_ptr, buf_len, int(stereo), buf.ctypes.data) < 0:
            raise RuntimeError("saugen_mix_to_pcm failed: " + last_error())
        return buf

    def op_state(self, op_id):
        v = OpView()
        if lib().saugen_read_op(self.ptr, op_id, C.byref(v)) != 0:
            raise IndexError(op_id)
        return v

    def voice_state(self, vo_id):
        out = (C.c_uint32 * 4)()
        lib().saugen_read_voice(self.ptr, vo_id, out)
        return list(out)

    def voice_rows(self, vo_id, n):
        s = np.zeros(n, np.float32)
        r = np.zeros(n, np.float32)
        if lib().saugen_read_voice_rows(self.ptr, vo_id, s.ctypes.data, r.ctypes.data, n) != 0:
            raise IndexError(vo_id)
        return s, r

    def debug_tap(self):
        """Developer aid: from now on keep every operator's output buffer of a call (general interpreter only)."""
        L = lib()
        L.saugen_debug_tap.restype = C.c_int
        L.saugen_debug_tap.argtypes = [C.c_void_p]
        if L.saugen_debug_tap(self.ptr) != 0:
            raise RuntimeError("saugen_debug_tap")

    def read_tap(self, op_id, n):
        L = lib()
        L.saugen_debug_read_tap.restype = C.c_int
        L.saugen_debug_read_tap.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        out = np.zeros(n, np.float32)
        if L.saugen_debug_read_tap(self.ptr, op_id, out.ctypes.data, n) != 0:
            raise IndexError(op_id)
        return out

    def counters(self):
        out = (C.c_uint64 * 4)()
        lib().saugen_counters(self.ptr, out)
        return list(out)

    def set_timing(self, on=True):
        lib().saugen_set_timing(self.ptr, int(on))

    def kernel_ms(self):
        """(render_kernel ms, mix_kernel ms) accumulated since set_timing(True)."""
        out = (C.c_double * 2)()
        lib().saugen_kernel_ms(self.ptr, out)
        return out[0], out[1]

    @property
    def amp_scale(self):
        return lib().saugen_amp_scale(self.ptr)

    def close(self):
        if getattr(self, "ptr", None):
            lib().saugen_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def voice_groups(prg, srate=96000):
    """saugen_voice_groups: (hand-over events, [group id per voice]); needs no GPU."""
    from . import program as P
    nvo = P.Program.from_address(prg.ptr).vo_count
    out = (C.c_uint32 * max(nvo, 1))()
    r = lib().saugen_voice_groups(prg.ptr, srate, out)
    if r < 0:
        raise RuntimeError("saugen_voice_groups failed: " + last_error())
    return r, list(out[:nvo])


def flatten(prg, srate=96000):
    """saugen_flatten: the program as one relocatable blob (bytes); needs no GPU.
    Generator(blob, ...) / render(blob, ...) instantiate it (the sample rate is the blob's)."""
    L = lib()
    need = L.saugen_flatten(prg.ptr, srate, None, 0)
    if not need:
        raise RuntimeError("saugen_flatten failed: " + last_error())
    buf = (C.c_char * need)()
    if L.saugen_flatten(prg.ptr, srate, buf, need) != need:
        raise RuntimeError("saugen_flatten failed: " + last_error())
    return bytes(buf)


class _DevArray:
    """Raw device memory exposed through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, numel, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr,
                                         "data": (ptr, False), "version": 2}


def planes_as_torch(ptr, numel):
    """View `numel` floats at device pointer `ptr` as a torch CUDA tensor (for the
    NCCL reduce of voice-sharded mixes); valid until the generator's next call."""
    import torch
    return torch.as_tensor(_DevArray(ptr, numel), device="cuda")


def render(prg, srate=96000, stereo=True, call_len=None, tables=None, device=0, max_frames=None,
           sched=0, big_endian=False):
    """Render a whole program the way Player_run does (saugns.c:575-623)."""
    if call_len is None:
        call_len = srate * 256 // 1000
    g = Generator(prg, srate, tables=tables, device=device, max_call_len=call_len, sched=sched,
                  big_endian=big_endian)
    ch = 2 if stereo else 1
    chunks, total, more = [], 0, True
    while more:
        more, buf, n = g.run(call_len, stereo)
        chunks.append(buf[:n * ch])
        total += n
        if max_frames and total >= max_frames:
            break
    g.close()
    return np.concatenate(chunks).reshape(-1, ch) if chunks else np.zeros((0, ch), np.int16)


def run_many(gens, buf_len, stereo=True, want_pcm=True):
    """One call on each generator with a single pair of kernel launches."""
    n = len(gens)
    ch = 2 if stereo else 1
    ptrs = (C.c_void_p * n)(*[g.ptr for g in gens])
    outs = np.zeros((n, buf_len * ch), np.int16) if want_pcm else None
    bufs = (C.c_void_p * n)(*[outs[i].ctypes.data for i in range(n)]) if want_pcm else None
    lens = (C.c_size_t * n)()
    more = (C.c_int * n)()
    r = lib().saugen_run_many(ptrs, n, bufs, buf_len, int(stereo), lens, more)
    if r < 0:
        raise RuntimeError("saugen_run_many failed: " + last_error())
    return list(more), outs, list(lens)
