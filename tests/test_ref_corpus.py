"""The reference's own script corpus (examples/**, devtests/**; 100 files,
96 of which its front end can build) on the CUDA path.  The GPU box has no /root/reference, so
the scripts travel as a fixture (tests/golden/ref_corpus.json, made by tests/golden/make_corpus.py).
-m gpu: every script through the C ABI, bit-exact against the unmodified reference generator
running beside it (oracle/_ref) and against the committed answer where this host's wave tables
are the ones it was made with; a subset through the UNMODIFIED reference CLI linked to the B200
back end, byte-identical WAV files.  CPU: the fixture is what the reference tree holds."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
with open(os.path.join(HERE, "golden", "ref_corpus.json")) as _f:
    CORPUS = json.load(_f)
BUILDS = [e for e in CORPUS["scripts"] if e["builds"]]
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "saugns_ref")
B200_CLI = os.path.join(ROOT, "oracle", "_ref", "saugns_b200_cli")
# through the command-line drop-in as well: the scripts VERDICT r01 names + one of each kind
CLI_SET = ["examples/rainy_thunder.sau", "examples/misc3-2pm_R.sau", "examples/sounds/pm_feedback_pm.sau",
           "devtests/voice-reuse.sau", "examples/tests/through-zero-morph.sau", "devtests/melody1-pm_vary.sau",
           "examples/tests/numexpr.sau", "examples/sounds/voicelike-Rcos_rm.sau", "devtests/pm-addremaddrem.sau",
           "examples/halfrect_ringmod.sau", "examples/tests/panning.sau", "examples/tests/line_noisy.sau"]


def test_fixture_matches_reference_tree():
    """Every .sau file of the reference tree is in the fixture, verbatim (CPU, where the tree exists)."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("no /root/reference here")
    import glob
    files = sorted(glob.glob(os.path.join(ref, "examples", "**", "*.sau"), recursive=True) +
                   glob.glob(os.path.join(ref, "devtests", "**", "*.sau"), recursive=True))
    assert [os.path.relpath(p, ref) for p in files] == [e["path"] for e in CORPUS["scripts"]]
    for p, e in zip(files, CORPUS["scripts"]):
        assert open(p).read() == e["text"], e["path"]
    assert len(BUILDS) >= 92
    assert all(e["path"].startswith("devtests/crashes/") or e["path"].startswith("devtests/warning/")
               for e in CORPUS["scripts"] if not e["builds"])


def test_fixture_answers_are_the_references(ref):
    """The committed answers are what the reference renders here (when the tables are the fixture's)."""
    t = ref.piluts()
    mine = {w: hashlib.sha256(t[i].tobytes()).hexdigest() for i, w in enumerate(ref.WAVES)}
    checked = 0
    for e in BUILDS:
        if e["frames"] > 3_000_000:
            continue
        if any((e["waves"] >> i) & 1 and CORPUS["_meta"]["tables"][w] != mine[w] for i, w in enumerate(ref.WAVES)):
            continue
        pcm = ref.render(ref.Program(e["text"]), srate=e["srate"])
        assert pcm.shape[0] == e["frames"], e["path"]
        assert hashlib.sha256(np.ascontiguousarray(pcm).tobytes()).hexdigest() == e["sha256"], e["path"]
        checked += 1
    assert checked >= 20


@pytest.fixture(scope="module")
def S():
    import saugns_b200
    if saugns_b200.device_count() < 1:
        pytest.fail("no CUDA device: the B200 back end has no CPU fallback")
    return saugns_b200


@pytest.fixture(scope="module")
def tabs(port):
    import gpuutil
    return gpuutil.ref_tables_for_gpu(port)


@pytest.mark.gpu
@pytest.mark.parametrize("e", BUILDS, ids=[e["path"] for e in BUILDS])
def test_corpus_script_bit_exact(S, ref, tabs, e):
    prg = ref.Program(e["text"])
    want = ref.render(prg, srate=e["srate"])
    got = S.render(prg, srate=e["srate"], tables=tabs)
    assert got.shape == want.shape, e["path"]
    assert np.array_equal(got, want), e["path"]
    t = ref.piluts()
    mine = {w: hashlib.sha256(t[i].tobytes()).hexdigest() for i, w in enumerate(ref.WAVES)}
    if all(not ((e["waves"] >> i) & 1) or CORPUS["_meta"]["tables"][w] == mine[w] for i, w in enumerate(ref.WAVES)):
        assert got.shape[0] == e["frames"]
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == e["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", CLI_SET)
def test_corpus_script_cli_wav_identical(path, tmp_path):
    """`saugns -o out.wav script.sau`: the stock reference binary and the unmodified CLI + front end
    linked against libsaugen_b200.so (oracle/Makefile `dropin`) write the same file."""
    if not (os.path.exists(REF_CLI) and os.path.exists(B200_CLI)):
        pytest.skip("oracle/_ref CLIs not built")
    e = next(x for x in BUILDS if x["path"] == path)
    src = tmp_path / os.path.basename(path)
    src.write_text(e["text"])
    outs = []
    for cli, name in ((REF_CLI, "a.wav"), (B200_CLI, "b.wav")):
        out = str(tmp_path / name)
        r = subprocess.run([cli, "-m", "-d", "-r", str(e["srate"]), "-o", out, str(src)],
                           capture_output=True, timeout=900)
        assert r.returncode == 0, (cli, r.stderr[-400:])
        with open(out, "rb") as f:
            outs.append(hashlib.sha256(f.read()).hexdigest())
        os.unlink(out)
    assert outs[0] == outs[1], path
