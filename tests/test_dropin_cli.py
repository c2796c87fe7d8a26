"""The drop-in, end to end (-m gpu): the reference's UNMODIFIED command-line
program and script front end (saugns.c, sau/parser.c ...), linked against
libsaugen_b200.so + dropin.o instead of sau/generator.o (oracle/Makefile target
`dropin`, the recipe of INTEGRATION.md), must write the same WAV file, byte
for byte, as the stock reference binary does: `saugns -o out.wav` on a B200.
Both binaries are prebuilt under oracle/_ref (they travel to the GPU box)."""
import os
import subprocess

import pytest

import scripts

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "saugns_ref")
B200_CLI = os.path.join(ROOT, "oracle", "_ref", "saugns_b200_cli")


def _render(cli, args, out):
    """`out` = a path (WAV file) or "-" (AU stream on stdout, saugns.c:508-511)."""
    r = subprocess.run([cli, "-m", "-d"] + args + ["-o", out], capture_output=True, timeout=600)
    assert r.returncode == 0, (cli, r.stderr)
    if out == "-":
        return r.stdout
    with open(out, "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def clis():
    if not (os.path.exists(REF_CLI) and os.path.exists(B200_CLI)):
        pytest.skip("oracle/_ref CLIs not built (run __graft_entry__.build() where /root/reference exists)")
    return REF_CLI, B200_CLI


def test_c1_wsin_wav_identical(clis, tmp_path):
    """BASELINE config 1: `saugns -e "Wsin"` -> 384 044-byte stereo 96 kHz WAV."""
    a = _render(clis[0], ["-r", "96000", "-e", "Wsin"], str(tmp_path / "a.wav"))
    b = _render(clis[1], ["-r", "96000", "-e", "Wsin"], str(tmp_path / "b.wav"))
    assert len(a) == 384044
    assert a == b


def test_c2_misc1_wav_identical(clis, tmp_path):
    """BASELINE config 2: examples/misc1-4fm_pm.sau rendered to WAV (script text
    carried in tests/scripts.py; the reference tree is absent on the GPU box)."""
    src = tmp_path / "misc1.sau"
    src.write_text(scripts.C2_MISC1_4FM_PM)
    a = _render(clis[0], ["-r", "96000", str(src)], str(tmp_path / "a.wav"))
    b = _render(clis[1], ["-r", "96000", str(src)], str(tmp_path / "b.wav"))
    assert len(a) == 23040044
    assert a == b


def test_several_scripts_one_output(clis, tmp_path):
    """Player_run with several scripts appended to one file (saugns.c:648-659),
    default 44.1 kHz, mixed operator types and voice reuse."""
    feats = scripts.feature_scripts()
    files = []
    for name in ["voices3", "seq_update", "self_w_mod", "noise_am", "R_cub_self", "pan_mod"]:
        p = tmp_path / (name + ".sau")
        p.write_text(feats[name] + "\n")
        files.append(str(p))
    a = _render(clis[0], files, str(tmp_path / "a.wav"))
    b = _render(clis[1], files, str(tmp_path / "b.wav"))
    assert len(a) > 44 and a == b


def test_mono_au_stdout_paths(clis, tmp_path):
    """--mono and the AU writer take the same PCM through other player code."""
    text = scripts.feature_scripts()["pm_chain"]
    for extra, outs in [(["--mono"], ("a.wav", "b.wav")), ([], ("-", "-"))]:
        oa, ob = [o if o == "-" else str(tmp_path / o) for o in outs]
        a = _render(clis[0], ["-r", "48000", "-e", text] + extra, oa)
        b = _render(clis[1], ["-r", "48000", "-e", text] + extra, ob)
        assert len(a) > 1000 and a == b
