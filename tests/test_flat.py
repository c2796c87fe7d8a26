"""The flat program (saugen_flatten / saugen_create_flat, include/saugen_b200.h): the
device-ready form of a sauProgram as one relocatable blob."""
import numpy as np
import pytest

import scripts


def test_flatten_needs_no_gpu_and_is_deterministic(ref):
    import saugns_b200
    feats = scripts.feature_scripts()
    for name in ["pm_chain", "fm_both", "seq_update", "voices3", "R_cub_self", "noise_am"]:
        prg = ref.Program(feats[name])
        a = saugns_b200.flatten(prg, 96000)
        b = saugns_b200.flatten(prg, 96000)
        assert a == b and len(a) > 64 and a[:4] == b"SAUF"
        assert saugns_b200.flatten(prg, 48000) != a          # times are in samples
    deep = "Wsin f200 t0.1 " + "p[Wsin r2 " * 40 + "]" * 40
    with pytest.raises(RuntimeError):                        # too deep for the device interpreter
        saugns_b200.flatten(ref.Program(deep), 96000)


def test_create_flat_rejects_garbage():
    import ctypes as C
    import saugns_b200
    L = saugns_b200.lib()
    junk = (C.c_char * 256)()
    assert not L.saugen_create_flat(junk, 256, None, None)
    assert b"flat program" in L.saugen_last_error()


@pytest.mark.gpu
def test_render_from_blob_equals_render_from_program(ref, port):
    """Flatten, drop the sauProgram, instantiate the blob: same PCM bit for bit, for the
    feature corpus and C5 scripts."""
    import gpuutil
    import saugns_b200
    tabs = gpuutil.ref_tables_for_gpu(port)
    texts = dict(scripts.feature_scripts())
    for i in range(6):
        texts[f"c5_{i}"] = scripts.synth_c5_script(i)
    bad = []
    for name, text in sorted(texts.items()):
        prg = ref.Program(text)
        want = saugns_b200.render(prg, srate=96000, tables=tabs)
        blob = saugns_b200.flatten(prg, 96000)
        del prg
        got = saugns_b200.render(blob, srate=96000, tables=tabs)
        if got.shape != want.shape or not np.array_equal(got, want):
            bad.append(name)
    assert not bad, bad


def test_operator_handover_between_voices_is_found(ref):
    """parseconv re-homes an operator whose voice slot was recycled; the flattener must see it
    (the run time cuts the call there, runtime.cpp:plan_call), and voice sharding must keep the
    linked voices together (saugen_voice_groups)."""
    import saugns_b200
    feats = scripts.feature_scripts()
    n, groups = saugns_b200.voice_groups(ref.Program(feats["handover"]))
    assert n == 1 and groups == [0, 0]
    n, groups = saugns_b200.voice_groups(ref.Program(feats["handover_same_time"]))
    assert n == 1 and groups == [0, 0]
    n, groups = saugns_b200.voice_groups(ref.Program(feats["handover_twice"]))
    assert n >= 2 and len(set(groups)) < len(groups)
    for name in ["voices3", "seq_overlap", "seq_update", "pm_addrem", "seq_bar"]:
        prg = ref.Program(feats[name])
        n, groups = saugns_b200.voice_groups(prg)
        assert n == 0 and groups == list(range(prg.vo_count)), name
    n, groups = saugns_b200.voice_groups(ref.Program(scripts.synth_c3(64, 1, fm="mix")))
    assert n == 0 and groups == list(range(64))
