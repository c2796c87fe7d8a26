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


def render_batch(programs, srate=96000, device=0, call_len=None, tables=None, group_size=256,
                 stereo=True, max_frames=None):
    """Render every program of `programs` -> list of int16 arrays [frames, ch],
    in input order.  Programs are independent (no mixing between them)."""
    if call_len is None:
        call_len = srate * 256 // 1000            # saugns.c:471
    ch = 2 if stereo else 1
    n = len(programs)
    out = [None] * n
    chunks = {}
    live = []                                     # [(index, Generator)]
    nxt = 0
    while nxt < n or live:
        while nxt < n and len(live) < group_size:
            g = G.Generator(programs[nxt], srate, tables=tables, device=device,
                            max_call_len=call_len)
            live.append((nxt, g))
            chunks[nxt] = []
            nxt += 1
        more, pcm, lens = G.run_many([g for _, g in live], call_len, stereo)
        keep = []
        for k, (i, g) in enumerate(live):
            if lens[k]:
                chunks[i].append(pcm[k][:lens[k] * ch].copy())
            done = not more[k]
            if max_frames and sum(c.size for c in chunks[i]) >= max_frames * ch:
                done = True
            if done:
                g.close()
                parts = chunks.pop(i)
                out[i] = (np.concatenate(parts).reshape(-1, ch) if parts
                          else np.zeros((0, ch), np.int16))
            else:
                keep.append((i, g))
        live = keep
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
