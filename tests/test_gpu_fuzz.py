"""Randomised parity (-m gpu): multi-voice scripts assembled from the feature corpus with
random delays, durations that put line goals and operator ends in the middle of calls and
of 1024-sample blocks, compared bit for bit (PCM, then integer and float operator state)
against the unmodified reference at several call sizes.  Aimed at the steady-stretch plan
(render_plan.cuh:steady_plan): stretches of many blocks next to blocks the general interpreter
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


def fuzz_sequence_script(seed):
    """Voices that are UPDATED while they run: `;` steps changing frequency (with and
    without sweeps), amplitude ramps, wave type, phase, modulator lists and pan -- every
    step is an event that ends a steady stretch, rewrites operator state and starts a new
    plan, at times that fall anywhere inside blocks and calls."""
    rnd = random.Random(10_000 + seed)
    waves = scripts.WAVES
    lines_ = scripts.LINES

    def mod():
        k = rnd.randrange(4)
        if k == 0:
            return f"p[W{rnd.choice(waves)} r{rnd.choice([0.5, 1, 2, 3])} a{rnd.uniform(0.1, 0.9):.3f}]"
        if k == 1:
            return f"p[W{rnd.choice(waves)} f{rnd.uniform(50, 900):.2f} a{rnd.uniform(0.1, 0.9):.3f}[g{rnd.uniform(0, 1):.2f} l{rnd.choice(lines_)}]]"
        if k == 2:
            return f"a{rnd.uniform(0.2, 1):.3f}.r{rnd.uniform(0, 0.5):.3f}[Wsin f{rnd.uniform(1, 12):.2f}]"
        return f"f{rnd.uniform(100, 800):.2f}.r{rnd.uniform(100, 1600):.2f}[W{rnd.choice(waves)} f{rnd.uniform(1, 30):.2f}]"

    def step():
        k = rnd.randrange(7)
        t = f"t{rnd.uniform(0.03, 0.4):.4f}"
        if k == 0:
            return f"f{rnd.uniform(60, 1500):.2f} {t}"
        if k == 1:
            return f"f[g{rnd.uniform(60, 1500):.2f} l{rnd.choice(lines_)} t{rnd.uniform(0.02, 0.5):.4f}] {t}"
        if k == 2:
            return f"a{rnd.uniform(0.1, 1):.3f}[g{rnd.uniform(0, 1):.3f} l{rnd.choice(lines_)} t{rnd.uniform(0.02, 0.5):.4f}] {t}"
        if k == 3:
            return f"w{rnd.choice(waves)} {t}"
        if k == 4:
            return f"p{rnd.uniform(0, 1):.3f} {t}"
        if k == 5:
            return f"c{rnd.uniform(-1, 1):.3f} {t}"
        return f"{mod()} {t}"

    voices = []
    for _ in range(rnd.randint(1, 4)):
        head = f"W{rnd.choice(waves)} f{rnd.uniform(80, 1200):.2f} t{rnd.uniform(0.05, 0.4):.4f} c{rnd.uniform(-1, 1):.3f}"
        if rnd.random() < 0.7:
            head += " " + mod()
        line = head + "".join("; " + step() for _ in range(rnd.randint(1, 5)))
        if voices and rnd.random() < 0.4:
            line = f"/{rnd.uniform(0.01, 0.3):.4f} " + line
        voices.append(line)
    return "\n".join(voices) + "\n"


@pytest.mark.parametrize("seed", range(120))
def test_random_update_sequences_bit_exact(ref, port, tabs, seed):
    import saugns_b200 as S
    text = fuzz_sequence_script(seed)
    prg = ref.Program(text)
    want = ref.render(prg, srate=96000)
    got = S.render(prg, srate=96000, tables=tabs, call_len=4 * 24576)
    assert got.shape == want.shape, (seed, text)
    assert np.array_equal(got, want), (seed, text)
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
        ncall += 1
