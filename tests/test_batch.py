"""Batched multi-program driver and sound-file bytes (saugns_b200/batch.py)."""
import os
import subprocess

import numpy as np
import pytest

import scripts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "saugns_ref")


@pytest.mark.parametrize("fmt,stereo", [("wav", True), ("wav", False), ("au", True)])
def test_sound_file_bytes_match_reference_cli(ref, tmp_path, fmt, stereo):
    """wav_bytes / au_bytes of the reference's PCM == the file the reference CLI writes
    (player/sndfile.c header layout, sizes patched on close, AU byte order)."""
    from saugns_b200 import batch
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/saugns_ref not built")
    text = scripts.feature_scripts()["voices3"]
    out = str(tmp_path / "x.wav")
    args = [REF_CLI, "-m", "-d", "-r", "48000", "-e", text, "-o", out if fmt == "wav" else "-"]
    if not stereo:
        args.insert(1, "--mono")
    r = subprocess.run(args, check=True, capture_output=True)
    if fmt == "au":                       # AU goes to stdout only (saugns.c:508-511)
        with open(out, "wb") as f:
            f.write(r.stdout)
    pcm = ref.render(ref.Program(text), srate=48000, stereo=stereo,
                     call_len=48000 * 256 // 1000)
    data = batch.wav_bytes(pcm, 48000) if fmt == "wav" else batch.au_bytes(pcm, 48000)
    assert data == open(out, "rb").read()


@pytest.mark.gpu
def test_render_batch_matches_reference(ref, port):
    """40 independent C5 scripts through saugen_run_many with admission /
    retirement between calls == each script rendered by the reference."""
    import gpuutil
    from saugns_b200 import batch
    tabs = gpuutil.ref_tables_for_gpu(port)
    prgs = [ref.Program(scripts.synth_c5_script(i)) for i in range(40)]
    got = batch.render_batch(prgs, srate=96000, tables=tabs, group_size=16)
    # several driver threads, longer calls, streaming sink with recycled arrays
    sunk = {}
    none = batch.render_batch(prgs, srate=96000, tables=tabs, group_size=8, threads=3,
                              call_len=4 * 24576, pinned=True,
                              sink=lambda i, pcm: sunk.__setitem__(i, pcm.copy()))
    assert none == [None] * len(prgs) and len(sunk) == len(prgs)
    for i, p in enumerate(prgs):
        want = ref.render(p, srate=96000)
        assert got[i].shape == want.shape, i
        assert np.array_equal(got[i], want), i
        assert np.array_equal(sunk[i], want), i


@pytest.mark.gpu
@pytest.mark.parametrize("stereo", [True, False])
def test_gpu_big_endian_epilogue_is_the_reference_au_stream(ref, port, tmp_path, stereo):
    """Rendered with pcm_big_endian (the byte swap of player/sndfile.c:160-168 folded into the
    mix epilogue) + the 28-byte header == the AU stream `saugns -o -` writes, byte for byte."""
    import gpuutil
    import saugns_b200
    from saugns_b200 import batch
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/saugns_ref not built")
    tabs = gpuutil.ref_tables_for_gpu(port)
    text = scripts.feature_scripts()["voices3"]
    args = [REF_CLI, "-m", "-d", "-r", "48000", "-e", text, "-o", "-"]
    if not stereo:
        args.insert(1, "--mono")
    want = subprocess.run(args, check=True, capture_output=True).stdout
    pcm = saugns_b200.render(ref.Program(text), srate=48000, stereo=stereo, tables=tabs,
                             call_len=48000 * 256 // 1000, big_endian=True)
    assert batch.au_bytes(pcm, 48000, swapped=True) == want
