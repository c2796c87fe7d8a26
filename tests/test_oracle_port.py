"""Pins the oracle: the scalar restatement (oracle/saugen_oracle.cpp, sharing
saugns_b200/csrc/sau_arith.h with the CUDA kernels) must reproduce the
UNMODIFIED reference generator (oracle/_ref/libsauref.so) bit for bit:
16-bit PCM, integer phase/cycle/counter state and float operator state.
CPU only."""
import glob
import os

import numpy as np
import pytest

import scripts

REFDIR = "/root/reference"


def _render_both(ref, port, prg, srate, stereo=True, call_len=None):
    a = ref.render(prg, srate=srate, stereo=stereo, call_len=call_len)
    b = port.render(prg, srate=srate, stereo=stereo, call_len=call_len)
    return a, b


@pytest.mark.parametrize("name,text", sorted(scripts.feature_scripts().items()))
def test_feature_script_pcm_bit_exact(ref, port, name, text):
    prg = ref.Program(text)
    a, b = _render_both(ref, port, prg, 96000)
    assert a.shape == b.shape
    assert a.shape[0] > 0
    assert np.array_equal(a, b)


def test_known_answer_c1(ref, port):
    """`-e "Wsin"` first left samples, frame count, final phase."""
    prg = ref.Program("Wsin")
    for mod in (ref, port):
        pcm = mod.render(prg, srate=96000)
        assert pcm.shape == (96000, 2)
        assert list(pcm[:8, 0]) == [447, 707, 1178, 1649, 2117, 2584, 3049, 3511]
        assert np.array_equal(pcm[:, 0], pcm[:, 1])
    g = port.PortGenerator(prg, 96000)
    more, _, n = g.run(24576)
    assert more and n == 24576
    assert g.op_state(0).i0 == 0x63d6c000
    while more:
        more, _, n = g.run(24576)
    assert g.op_state(0).i0 == 0xbffede00


@pytest.mark.parametrize("call_len", [24576, 1024, 1000, 333, 7])
def test_call_size_invariance_and_state(ref, port, call_len):
    """State after every call must agree bit for bit, for any call size."""
    names = ["pm_chain", "fm_both", "self_w_mod", "self_r_pm", "seq_update", "voices3",
             "noise_am", "R_cub_self", "sweep_f_cub", "regoal", "pan_mod", "mod_finite"]
    feats = scripts.feature_scripts()
    for name in names:
        prg = ref.Program(feats[name])
        gr = ref.RefGenerator(prg, 48000)
        gp = port.PortGenerator(prg, 48000)
        more = True
        ncall = 0
        while more and ncall < 400:
            more, ba, na = gr.run(call_len)
            more2, bb, nb = gp.run(call_len)
            assert (more, na) == (more2, nb), name
            assert np.array_equal(ba, bb), name
            for op in range(prg.op_count):
                assert port.op_state_tuple(gr.op_state(op)) == port.op_state_tuple(gp.op_state(op)), (name, op)
            for vo in range(prg.vo_count):
                assert gr.voice_state(vo) == gp.voice_state(vo)
            ncall += 1


def test_mono(ref, port):
    prg = ref.Program(scripts.feature_scripts()["voices3"])
    a, b = _render_both(ref, port, prg, 44100, stereo=False)
    assert a.shape[1] == 1 and np.array_equal(a, b)


@pytest.mark.skipif(not os.path.isdir(REFDIR), reason="reference tree absent")
def test_reference_example_scripts(ref, port):
    """All scripts shipped with the reference (examples/, devtests/)."""
    files = sorted(glob.glob(REFDIR + "/examples/*.sau") + glob.glob(REFDIR + "/examples/*/*.sau")
                   + glob.glob(REFDIR + "/devtests/*.sau"))
    assert len(files) >= 80
    for f in files:
        prg = ref.Program(f, is_path=True)
        a, b = _render_both(ref, port, prg, 24000)
        assert a.shape == b.shape, f
        assert np.array_equal(a, b), f


def test_synthetic_configs_small(ref, port):
    for text in (scripts.synth_c3(16, 1), scripts.synth_c3(8, 1, fm=True), scripts.synth_c4(12, 1),
                 scripts.synth_c5_script(0), scripts.synth_c5_script(7)):
        prg = ref.Program(text)
        a, b = _render_both(ref, port, prg, 96000)
        assert a.shape == b.shape and a.shape[0] > 0
        assert np.array_equal(a, b)


def test_float_buffers_match(ref, port):
    """Carrier float buffer of the last block (gen_bufs[0]) is bit-identical."""
    for name in ["pm_chain", "fm_range", "self_w", "R_xpe_perlin"]:
        prg = ref.Program(scripts.feature_scripts()[name])
        gr = ref.RefGenerator(prg, 96000)
        gp = port.PortGenerator(prg, 96000)
        for _ in range(5):
            gr.run(1024)
            gp.run(1024)
            assert np.array_equal(gr.gen_buf(0).view(np.uint32), gp.gen_buf(0).view(np.uint32)), name


def test_port_reproduces_golden_answers(ref, port):
    """The oracle port against the committed answers of the unmodified reference
    (tests/golden/, frames + sha256 of PCM + final integer state), wherever this
    host's wave tables are the ones the answers were made with."""
    import gpuutil
    gold = gpuutil.golden(ref)
    feats = scripts.feature_scripts()
    cases = [("feat/" + k, v) for k, v in sorted(feats.items())]
    cases += [("config/C1_Wsin", "Wsin"), ("config/C3_64v_1s", scripts.synth_c3(64, 1)),
              ("config/C4_48v_1s", scripts.synth_c4(48, 1))]
    cases += [(f"config/C5_script{i}", scripts.synth_c5_script(i)) for i in range(8)]
    checked = 0
    for key, text in cases:
        g = gold.get(key)
        if g is None:
            continue
        prg = ref.Program(text)
        pcm = port.render(prg, srate=g["srate"], stereo=g["stereo"])
        assert pcm.shape[0] == g["frames"], key
        assert gpuutil.sha(pcm) == g["sha256"], key
        checked += 1
    assert checked >= 150


def test_sweeps_past_2_24_samples(ref, port):
    """Line positions beyond 2**24 samples of one sweep (SURVEY.md App. B.1's untested note): the
    port's single conversion equals the as-compiled two-halves conversion."""
    prg = ref.Program(scripts.long_sweep_script())
    a = ref.render(prg, srate=96000)
    assert a.shape[0] > (1 << 24)
    assert np.array_equal(a, port.render(prg, srate=96000))
