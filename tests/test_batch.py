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


@pytest.mark.gpu
def test_native_batch_driver_matches_reference(ref, port):
    """saugen_render_batch (csrc/batch_driver.cpp): the batch loop with no Python between the calls."""
    import gpuutil
    from saugns_b200 import batch
    tabs = gpuutil.ref_tables_for_gpu(port)
    texts = [scripts.synth_c5_script(i) for i in range(40)] + [scripts.feature_scripts()[k] for k in
                                                               ("voices3", "seq_overlap", "handover_twice", "self_w_mod")]
    prgs = [ref.Program(t) for t in texts]
    got = batch.render_batch_native(prgs, srate=96000, tables=tabs, group_size=16)
    for i, p in enumerate(prgs):
        want = ref.render(p, srate=96000)
        assert got[i].shape == want.shape and np.array_equal(got[i], want), i
    seen = {}
    batch.render_batch_native(prgs[:12], srate=48000, tables=tabs, group_size=5, stereo=False,
                              sink=lambda i, pcm: seen.__setitem__(i, pcm.copy()))
    for i in range(12):
        assert np.array_equal(seen[i], ref.render(prgs[i], srate=48000, stereo=False)), i


@pytest.mark.gpu
def test_batch_cli_writes_the_references_wav_files(tmp_path):
    """`saugns_b200_batch -o dir a.sau b.sau ...` (the reference's own front end + the native batch
    driver + its WAV writer threads) == `saugns -m -d -o x.wav x.sau` of the stock reference, per file."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_cli = os.path.join(root, "oracle", "_ref", "saugns_ref")
    cli = os.path.join(root, "oracle", "_ref", "saugns_b200_batch")
    if not (os.path.exists(ref_cli) and os.path.exists(cli)):
        pytest.skip("oracle/_ref CLIs not built")
    files = []
    for i in range(24):
        p = tmp_path / f"s{i}.sau"
        p.write_text(scripts.synth_c5_script(100 + i))
        files.append(str(p))
    out = tmp_path / "out"
    out.mkdir()
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(files) + "\n")
    r = subprocess.run([cli, "-r", "96000", "-o", str(out), "-l", str(lst)], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    for i, f in enumerate(files):
        a = tmp_path / f"ref{i}.wav"
        rr = subprocess.run([ref_cli, "-m", "-d", "-r", "96000", "-o", str(a), f], capture_output=True, timeout=600)
        assert rr.returncode == 0
        assert a.read_bytes() == (out / f"s{i}.wav").read_bytes(), i
