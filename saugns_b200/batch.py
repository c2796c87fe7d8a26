"""Batched multi-program driver + sound-file output (SURVEY.md section 8f,
rank 1): what `Player_run` (saugns.c:575-665) and `player/sndfile.c` do for
one script at a time, for thousands of independent programs on one GPU.

`render_batch` keeps live sets of `group_size` generators and advances each set
with ONE render + ONE mix launch per call (`saugen_batch_begin` / `_end`, the
two halves of `saugen_run_many`), retiring finished programs and admitting new
ones between calls, while the other set's kernels run.  `write_wav` /
`write_au` keep the reference's byte format (player/sndfile.c:63-109): 44-byte
RIFF/WAVE header with sizes patched at close, little-endian int16; AU (the
`-o -` stdout stream) = 28-byte header, size unspecified, big-endian int16.
"""
import os
import struct

import numpy as np

from . import generator as G


def _capacity_frames(prg, srate, call_len):
    """Upper bound of the frames a program renders (its duration rounded up to whole calls)."""
    from . import program as P
    ms = P.Program.from_address(prg.ptr).duration_ms
    frames = (ms * srate + 999) // 1000
    return (frames // call_len + 2) * call_len


def render_batch(programs, srate=96000, device=0, call_len=None, tables=None, group_size=128,
                 stereo=True, max_frames=None, threads=1, sink=None, depth=2, pinned=False):
    """Render every program of `programs` -> list of int16 arrays [frames, ch],
    in input order.  Programs are independent (no mixing between them).

    A driver thread alternates `depth` live sets of up to `group_size` generators
    each: while the kernels of one set's call run (`saugen_batch_begin`), the host
    finishes the other set's previous call (`saugen_batch_end`), retires finished
    programs, admits new ones from the shared queue and plans the next call.
    Each call's PCM goes from the generator's page-locked staging block to the
    program's own array inside `saugen_batch_end` (a few copy threads), or, with
    `pinned=True`, the arrays themselves are page-locked (`saugen_pinned_alloc`)
    and the device-to-host copy is the only copy -- worth it when the arrays are
    recycled for long (page-locking memory costs about as much as copying into
    it a few times).  `threads` > 1 runs several such drivers (each with its own
    sets and streams).

    `sink(index, pcm)`: called (on a driver thread) with each finished program's
    PCM instead of keeping it; the array is only valid during the call -- its
    memory goes back to a pool and carries a later program's samples.  This is
    the streaming form (what Player_run does with its 256 ms buffer, saugns.c:
    601-609: write it out, reuse it): a batch of thousands of scripts does not
    hold gigabytes of PCM.  The returned list then holds None."""
    import itertools
    import threading
    if call_len is None:
        call_len = srate * 256 // 1000            # saugns.c:471
    n = len(programs)
    out = [None] * n
    queue = itertools.count()                     # next program index; next() is atomic under the GIL
    threads = max(1, min(int(threads), (n + 15) // 16))
    args = (programs, out, queue, srate, device, call_len, tables, group_size, stereo, max_frames,
            sink, max(1, int(depth)), bool(pinned))
    if threads == 1:
        _render_worker(*args)
        return out
    errs = []

    def work():
        try:
            _render_worker(*args)
        except BaseException as e:                # re-raised on the caller's thread
            errs.append(e)

    th = [threading.Thread(target=work) for _ in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return out


class _ArrayPool:
    """Recycled int16 output arrays: pageable numpy memory, or page-locked memory
    from the library's pool."""

    def __init__(self, pinned):
        self.free = []                            # (array, address)
        self.pinned = pinned
        self.L = G.lib()

    def take(self, size):
        for j, (arr, addr) in enumerate(self.free):
            if arr.size >= size:
                return self.free.pop(j)
        if not self.pinned:
            arr = np.empty(size, np.int16)
            return arr, arr.ctypes.data
        import ctypes as C
        size = 1 << (size - 1).bit_length()       # few distinct sizes: they recycle
        addr = self.L.saugen_pinned_alloc(size * 2)
        if not addr:
            raise MemoryError("saugen_pinned_alloc failed: " + G.last_error())
        arr = np.ctypeslib.as_array((C.c_int16 * size).from_address(addr))
        return arr, addr

    def give(self, arr, addr):
        self.free.append((arr, addr))

    def close(self):
        if self.pinned:
            for _, addr in self.free:
                self.L.saugen_pinned_free(addr)
        self.free = []


class _LiveSet:
    """Generators advanced together by one saugen_batch_begin / _end pair per call."""

    def __init__(self, device):
        self.L = G.lib()
        self.batch = self.L.saugen_batch_create(device)
        if not self.batch:
            raise RuntimeError("saugen_batch_create failed: " + G.last_error())
        self.idx, self.gens, self.bufs = [], [], []   # parallel lists; bufs = (array, address)
        self.gptr = np.zeros(0, np.uint64)            # generator handles
        self.base = np.zeros(0, np.uint64)            # address of each output array
        self.pos = np.zeros(0, np.int64)              # frames written so far
        self.cap = np.zeros(0, np.int64)
        self.in_flight = False

    def close(self):
        if self.batch:
            self.L.saugen_batch_destroy(self.batch)
            self.batch = None


def _render_worker(programs, out, queue, srate, device, call_len, tables, group_size, stereo,
                   max_frames, sink, depth, pinned):
    """One driver: `depth` live sets, admitted from `queue`, alternated call by call."""
    L = G.lib()
    ch = 2 if stereo else 1
    n = len(programs)
    pool = _ArrayPool(pinned)
    sets = [_LiveSet(device) for _ in range(depth)]
    drained = False

    def admit(s):
        nonlocal drained
        if drained or len(s.gens) >= group_size:
            return
        k0 = len(s.gens)
        while len(s.gens) < group_size:
            i = next(queue)
            if i >= n:
                drained = True
                break
            g = G.Generator(programs[i], srate, tables=tables, device=device, max_call_len=call_len)
            c = _capacity_frames(programs[i], srate, call_len)
            s.idx.append(i); s.gens.append(g); s.bufs.append(pool.take(c * ch))
        new = s.bufs[k0:]
        s.gptr = np.concatenate([s.gptr, np.array([g.ptr for g in s.gens[k0:]], np.uint64)])
        s.base = np.concatenate([s.base, np.array([a for _, a in new], np.uint64)])
        s.pos = np.concatenate([s.pos, np.zeros(len(new), np.int64)])
        s.cap = np.concatenate([s.cap, np.array([b.size // ch for b, _ in new], np.int64)])

    def begin(s):
        for k in np.nonzero(s.pos + call_len > s.cap)[0]:      # rare: longer than announced
            old, old_addr = s.bufs[k]
            arr, addr = pool.take(old.size + 4 * call_len * ch)
            arr[:s.pos[k] * ch] = old[:s.pos[k] * ch]
            pool.give(old, old_addr)
            s.bufs[k] = (arr, addr)
            s.base[k] = addr
            s.cap[k] = arr.size // ch
        s.ptrs = s.base + (s.pos * (2 * ch)).astype(np.uint64)   # kept alive until end()
        r = L.saugen_batch_begin(s.batch, s.gptr.ctypes.data, len(s.gens), s.ptrs.ctypes.data,
                                 call_len, int(stereo), int(pinned))
        if r < 0:
            raise RuntimeError("saugen_batch_begin failed: " + G.last_error())
        s.in_flight = True

    def end(s):
        m = len(s.gens)
        lens = np.zeros(m, np.uint64)
        more = np.zeros(m, np.int32)
        r = L.saugen_batch_end(s.batch, lens.ctypes.data, more.ctypes.data)
        s.in_flight = False
        if r < 0:
            raise RuntimeError("saugen_batch_end failed: " + G.last_error())
        s.pos += lens.astype(np.int64)
        done = more == 0
        if max_frames:
            done |= s.pos >= max_frames
        if not done.any():
            return
        keep = []
        for k in range(m):
            if not done[k]:
                keep.append(k)
                continue
            s.gens[k].close()
            arr, addr = s.bufs[k]
            pcm = arr[:s.pos[k] * ch].reshape(-1, ch)
            if sink is not None:
                sink(s.idx[k], pcm)
                pool.give(arr, addr)
            elif pinned:
                out[s.idx[k]] = pcm.copy()        # out of the page-locked pool
                pool.give(arr, addr)
            else:
                out[s.idx[k]] = pcm               # the array is the result
        s.idx = [s.idx[k] for k in keep]; s.gens = [s.gens[k] for k in keep]
        s.bufs = [s.bufs[k] for k in keep]
        kk = np.array(keep, np.int64)
        s.gptr, s.base, s.pos, s.cap = s.gptr[kk], s.base[kk], s.pos[kk], s.cap[kk]

    try:
        k = 0
        while True:
            s = sets[k % depth]
            k += 1
            if s.in_flight:
                end(s)
            admit(s)
            if s.gens:
                begin(s)
            if drained and not any(x.in_flight for x in sets):
                break
    finally:
        for s in sets:
            if s.in_flight:
                try:
                    end(s)
                except Exception:
                    pass
            for g in s.gens:
                g.close()
            s.close()
        pool.close()


class BatchOptions(__import__("ctypes").Structure):
    """saugen_BatchOptions (include/saugen_b200.h)."""
    _fields_ = [("device", __import__("ctypes").c_int), ("call_len", __import__("ctypes").c_uint32),
                ("group", __import__("ctypes").c_uint32), ("depth", __import__("ctypes").c_uint32),
                ("mono", __import__("ctypes").c_uint32), ("io_threads", __import__("ctypes").c_uint32)]


def render_batch_native(programs, srate=96000, device=0, call_len=0, tables=None, group_size=128, depth=2,
                        stereo=True, sink=None, wav_paths=None, io_threads=4, discard=False):
    """The same batch through the NATIVE driver (saugen_render_batch / saugen_render_batch_wav,
    csrc/batch_driver.cpp): no Python between the calls.  `wav_paths`: program i is written to
    wav_paths[i] by the library's writer threads (the reference's WAV format); else `sink(index,
    pcm)` gets every finished program's int16 array [frames, ch] (valid during the call) on the
    driver thread, or, with neither, the arrays are returned as a list (copies)."""
    import ctypes as C
    L = G.lib()
    if tables is None:
        tables = G.default_tables()
    n = len(programs)
    ptrs = (C.c_void_p * max(n, 1))(*[p.ptr for p in programs])
    opt = BatchOptions(device=device, call_len=call_len, group=group_size, depth=depth,
                       mono=0 if stereo else 1, io_threads=io_threads)
    if wav_paths is not None:
        keep = [os.fsencode(p) for p in wav_paths]
        arr = (C.c_char_p * max(n, 1))(*keep)
        r = L.saugen_render_batch_wav(ptrs, n, srate, C.addressof(tables), C.byref(opt), arr)
        if r < 0:
            raise RuntimeError("saugen_render_batch_wav failed: " + L.saugen_batch_last_error().decode())
        return None
    if discard:                # every script's PCM reaches host memory and is dropped there
        r = L.saugen_render_batch(ptrs, n, srate, C.addressof(tables), C.byref(opt), None, None)
        if r < 0:
            raise RuntimeError("saugen_render_batch failed: " + L.saugen_batch_last_error().decode())
        return None
    out = [None] * n
    ch = 2 if stereo else 1
    errs = []

    @C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int)
    def cb(user, index, pcm, frames, channels):
        try:
            a = np.ctypeslib.as_array((C.c_int16 * (frames * channels)).from_address(pcm)).reshape(-1, channels) \
                if frames else np.zeros((0, ch), np.int16)
            if sink is not None:
                sink(index, a)
            else:
                out[index] = a.copy()
        except BaseException as e:        # never unwind through the C driver
            errs.append(e)
        finally:
            L.saugen_pinned_free(pcm)

    r = L.saugen_render_batch(ptrs, n, srate, C.addressof(tables), C.byref(opt), cb, None)
    if errs:
        raise errs[0]
    if r < 0:
        raise RuntimeError("saugen_render_batch failed: " + L.saugen_batch_last_error().decode())
    return out


def wav_bytes(pcm, srate):
    """The bytes `saugns -o x.wav` writes for this PCM (player/sndfile.c:83-109)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    ch = pcm.shape[1] if pcm.ndim == 2 else 1
    data = pcm.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, srate, ch * srate * 2, ch * 2, 16) + b"data" + struct.pack(
        "<I", len(data))
    return hdr + data


def au_bytes(pcm, srate, swapped=False):
    """The AU stream `saugns -o -` writes to stdout (saugns.c:508-511,
    player/sndfile.c:63-72,160-168): size field left "unspecified" because a
    stream is never patched (sndfile.c:201-211), big-endian samples.
    swapped: `pcm` was rendered with big_endian=True (the byte swap already
    happened in the mix epilogue on the GPU) and goes out as it is."""
    ch = pcm.shape[1] if pcm.ndim == 2 else 1
    hdr = b".snd" + struct.pack(">IIIII", 28, 0xffffffff, 3, srate, ch) + struct.pack(">I", 0)
    if swapped:
        return hdr + np.ascontiguousarray(pcm).tobytes()
    return hdr + np.ascontiguousarray(pcm, dtype="<i2").astype(">i2").tobytes()


def write_wav(path, pcm, srate):
    with open(path, "wb") as f:
        f.write(wav_bytes(pcm, srate))


def write_au(path, pcm, srate, swapped=False):
    with open(path, "wb") as f:
        f.write(au_bytes(pcm, srate, swapped))
