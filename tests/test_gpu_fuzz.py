"""Randomised parity (-m gpu): multi-voice scripts assembled from the feature corpus with
random delays, durations that put line goals and operator ends in the middle of calls and
of 1024-sample blocks, compared bit for bit (PCM, then integer and float operator state)
against the unmodified reference at several call sizes.  Aimed at the steady-stretch plan
(kernels.cu:steady_plan): stretches of many blocks next to blocks the general interpreter
has to take, uniform-frequency ratio chains, amplitude modulators, N / R operators."""
import random

import numpy as np
import pytest

import gpuutil
import scripts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tabs(port):
    return gpuutil.ref_tables_for_gpu(port)


def fuzz_script(seed):
    rnd = random.Random(seed)
    feats = scripts.feature_scripts()
    # single-line voices only (sequences with ';' or '|' restructure the script around them)
    pool = sorted(k for k, v in feats.items() if ";" not in v and "|" not in v and not v.startswith("S "))
    lines = []
    for _ in range(rnd.randint(2, 6)):
        text = feats[rnd.choice(pool)]
        if rnd.random() < 0.5:                      # stretch or squeeze every duration in the line
            k = rnd.choice([0.37, 0.81, 1.9, 3.3, 6.7])
            out, i = [], 0
            while i < len(text):
                if text[i] == "t" and i + 1 < len(text) and (text[i + 1].isdigit() or text[i + 1] == "."):
                    j = i + 1
                    while j < len(text) and (text[j].isdigit() or text[j] == "."):
                        j += 1
                    out.append(f"t{float(text[i + 1:j]) * k:.4f}")
                    i = j
                else:
                    out.append(text[i])
                    i += 1
            text = "".join(out)
        if lines and rnd.random() < 0.4:
            text = f"/{rnd.uniform(0.01, 0.35):.4f} " + text
        lines.append(text)
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(150))
def test_random_multivoice_scripts_bit_exact(ref, port, tabs, seed):
    import saugns_b200 as S
    text = fuzz_script(seed)
    prg = ref.Program(text)
    want = ref.render(prg, srate=96000)
    for call_len in (4 * 24576, 5000):
        got = S.render(prg, srate=96000, tables=tabs, call_len=call_len)
        assert got.shape == want.shape, (seed, call_len, text)
        assert np.array_equal(got, want), (seed, call_len, text)
    # the reference player's call size: PCM and ALL operator / voice state after every call
    call_len = 24576
    gr = ref.RefGenerator(prg, 96000)
    gg = S.Generator(prg, 96000, tables=tabs, max_call_len=call_len)
    more, ncall = True, 0
    while more and ncall < 100:
        more, ba, na = gr.run(call_len)
        more2, bb, nb = gg.run(call_len)
        assert (more, na) == (more2, nb), (seed, ncall, text)
        assert np.array_equal(ba, bb), (seed, ncall, text)
        for op in range(prg.op_count):
            a = port.op_state_tuple(gr.op_state(op))
            b = port.op_state_tuple(gg.op_state(op))
            assert a == b, (seed, ncall, op, text)
        for vo in range(prg.vo_count):
            assert gr.voice_state(vo)[:3] == gg.voice_state(vo)[:3], (seed, ncall, vo, text)
        ncall += 1
