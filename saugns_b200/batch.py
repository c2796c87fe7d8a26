"""Batched multi-program driver + sound-file output (SURVEY.md section 8f,
rank 1): what `Player_run` (saugns.c:575-665) and `player/sndfile.c` do for
one script at a time, for thousands of independent programs on one GPU.

`render_batch` keeps `group_size` generators in flight and advances all of them
with ONE render + ONE mix launch per 256 ms call (`saugen_run_many`), retiring
finished programs and admitting new ones between calls.  `write_wav` /
`write_au` keep the reference's byte format (player/sndfile.c:63-109): 44-byte
RIFF/WAVE header with sizes patched at close, little-endian int16; AU (the
`-o -` stdout stream) = 28-byte header, size unspecified, big-endian int16.
"""
import struct

import numpy as np

from . import generator as G


def _capacity_frames(prg, srate, call_len):
    """Upper bound of the frames a program renders (its duration rounded up to whole calls)."""
    from . import program as P
    ms = P.Program.from_address(prg.ptr).duration_ms
    frames = (ms * srate + 999) // 1000
    return (frames // call_len + 2) * call_len


def render_batch(programs, srate=96000, device=0, call_len=None, tables=None, group_size=256,
                 stereo=True, max_frames=None, threads=1, sink=None):
    """Render every program of `programs` -> list of int16 arrays [frames, ch],
    in input order.  Programs are independent (no mixing between them).

    `threads` host threads each keep their own live set of up to `group_size`
    generators (admitted from one shared queue) on their own stream: while one
    set's kernels run, the other threads plan calls, create and retire generators
    and copy PCM (all of that happens inside the C library, without the GIL), so
    the GPU does not wait for the host between calls.

    Each call's PCM lands directly in the program's final array (the C side
    copies from its pinned staging buffer to the address it is given), so the
    per-call Python work is a few vector operations over the live set.

    `sink(index, pcm)`: called (on a worker thread) with each finished program's
    PCM instead of keeping it; the array is only valid during the call -- its
    memory goes back to a pool and carries a later program's samples.  This is
    the streaming form (what Player_run does with its 256 ms buffer, saugns.c:
    601-609: write it out, reuse it): a batch of thousands of scripts does not
    hold gigabytes of PCM, nor pay the first-touch page faults of fresh arrays.
    The returned list then holds None."""
    import itertools
    import threading
    if call_len is None:
        call_len = srate * 256 // 1000            # saugns.c:471
    n = len(programs)
    out = [None] * n
    queue = itertools.count()                     # next program index; next() is atomic under the GIL
    threads = max(1, min(int(threads), (n + 15) // 16))
    if threads == 1:
        _render_worker(programs, out, queue, srate, device, call_len, tables, group_size, stereo,
                       max_frames, sink)
        return out
    errs = []

    def work():
        try:
            _render_worker(programs, out, queue, srate, device, call_len, tables, group_size,
                           stereo, max_frames, sink)
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


def _render_worker(programs, out, queue, srate, device, call_len, tables, group_size, stereo,
                   max_frames, sink=None):
    """One live set: admit from `queue`, advance with saugen_run_many, retire."""
    L = G.lib()
    ch = 2 if stereo else 1
    n = len(programs)
    idx, gens, bufs = [], [], []                  # the live set, parallel lists
    gptr = np.zeros(0, np.uint64)                 # generator handles
    base = np.zeros(0, np.uint64)                 # address of each output array
    pos = np.zeros(0, np.int64)                   # frames written so far
    cap = np.zeros(0, np.int64)
    drained = False
    pool = []                                     # recycled output arrays (sink mode)

    def new_buf(size):
        for j, b in enumerate(pool):
            if b.size >= size:
                return pool.pop(j)
        return np.empty(size, np.int16)

    while not drained or gens:
        if not drained and len(gens) < group_size:
            k0 = len(gens)
            while len(gens) < group_size:
                i = next(queue)
                if i >= n:
                    drained = True
                    break
                g = G.Generator(programs[i], srate, tables=tables, device=device,
                                max_call_len=call_len)
                c = _capacity_frames(programs[i], srate, call_len)
                b = new_buf(c * ch)
                idx.append(i); gens.append(g); bufs.append(b)
            gptr = np.concatenate([gptr, np.array([g.ptr for g in gens[k0:]], np.uint64)])
            base = np.concatenate([base, np.array([b.ctypes.data for b in bufs[k0:]], np.uint64)])
            pos = np.concatenate([pos, np.zeros(len(gens) - k0, np.int64)])
            cap = np.concatenate([cap, np.array([b.size // ch for b in bufs[k0:]], np.int64)])
        m = len(gens)
        if m == 0:
            break
        for k in np.nonzero(pos + call_len > cap)[0]:      # rare: longer than announced
            bufs[k] = np.concatenate([bufs[k], np.empty(4 * call_len * ch, np.int16)])
            base[k] = bufs[k].ctypes.data
            cap[k] = bufs[k].size // ch
        ptrs = base + (pos * (2 * ch)).astype(np.uint64)
        lens = np.zeros(m, np.uint64)
        more = np.zeros(m, np.int32)
        r = L.saugen_run_many(gptr.ctypes.data, m, ptrs.ctypes.data, call_len, int(stereo),
                              lens.ctypes.data, more.ctypes.data)
        if r < 0:
            raise RuntimeError("saugen_run_many failed: " + G.last_error())
        pos += lens.astype(np.int64)
        done = more == 0
        if max_frames:
            done |= pos >= max_frames
        if done.any():
            keep = []
            for k in range(m):
                if done[k]:
                    gens[k].close()
                    pcm = bufs[k][:pos[k] * ch].reshape(-1, ch)
                    if sink is None:
                        out[idx[k]] = pcm
                    else:
                        sink(idx[k], pcm)
                        pool.append(bufs[k])
                else:
                    keep.append(k)
            idx = [idx[k] for k in keep]; gens = [gens[k] for k in keep]; bufs = [bufs[k] for k in keep]
            kk = np.array(keep, np.int64)
            gptr, base, pos, cap = gptr[kk], base[kk], pos[kk], cap[kk]


def wav_bytes(pcm, srate):
    """The bytes `saugns -o x.wav` writes for this PCM (player/sndfile.c:83-109)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    ch = pcm.shape[1] if pcm.ndim == 2 else 1
    data = pcm.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, srate, ch * srate * 2, ch * 2, 16) + b"data" + struct.pack(
        "<I", len(data))
    return hdr + data


def au_bytes(pcm, srate):
    """The AU stream `saugns -o -` writes to stdout (saugns.c:508-511,
    player/sndfile.c:63-72,160-168): size field left "unspecified" because a
    stream is never patched (sndfile.c:201-211), big-endian samples."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    ch = pcm.shape[1] if pcm.ndim == 2 else 1
    hdr = b".snd" + struct.pack(">IIIII", 28, 0xffffffff, 3, srate, ch) + struct.pack(">I", 0)
    return hdr + pcm.astype(">i2").tobytes()


def write_wav(path, pcm, srate):
    with open(path, "wb") as f:
        f.write(wav_bytes(pcm, srate))


def write_au(path, pcm, srate):
    with open(path, "wb") as f:
        f.write(au_bytes(pcm, srate))
