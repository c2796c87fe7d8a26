"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI
(libsaugen_b200.so), against the unmodified reference (oracle/_ref), the
scalar port, and the committed golden answers.  Bar: bit-exact PCM and
integer state (stricter than the +/-1 LSB the north star allows); float
oscillator rows within 1e-5 relative (observed: identical bits)."""
import os

import numpy as np
import pytest

import gpuutil
import scripts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import saugns_b200
    if saugns_b200.device_count() < 1:
        pytest.fail("no CUDA device: the B200 back end has no CPU fallback")
    return saugns_b200


@pytest.fixture(scope="module")
def tabs(port):
    return gpuutil.ref_tables_for_gpu(port)


@pytest.fixture(scope="module")
def gold():
    return gpuutil.golden()


def test_fast_path_arithmetic_selftest(S, tabs):
    """Hand-expanded division / rounding primitives == the plain IEEE statements."""
    from saugns_b200 import generator
    assert generator.selftest(0, tabs) == 0
    assert generator.selftest(0, None) == 0


def test_feature_scripts_bit_exact(S, ref, tabs, gold):
    bad = []
    for name, text in sorted(scripts.feature_scripts().items()):
        prg = ref.Program(text)
        want = ref.render(prg, srate=96000)
        got = S.render(prg, srate=96000, tables=tabs)
        g = gold.get("feat/" + name)
        if got.shape != want.shape or not np.array_equal(got, want):
            bad.append(name)
        elif g is not None and (got.shape[0] != g["frames"] or gpuutil.sha(got) != g["sha256"]):
            bad.append(name + "(golden)")
    assert not bad, bad
    assert gold.applied >= 150      # the table-independent answers always apply


def test_c1_known_answer(S, ref, tabs, gold):
    prg = ref.Program("Wsin")
    g = S.Generator(prg, 96000, tables=tabs)
    more, buf, n = g.run(24576)
    assert more and n == 24576
    assert list(buf[0:16:2]) == [447, 707, 1178, 1649, 2117, 2584, 3049, 3511]
    assert g.op_state(0).i0 == 0x63d6c000
    chunks = [buf]
    while more:
        more, buf, n = g.run(24576)
        chunks.append(buf[:2 * n])
    assert g.op_state(0).i0 == 0xbffede00
    pcm = np.concatenate(chunks)
    assert pcm.size == 2 * 96000
    assert gpuutil.sha(pcm) == gold["config/C1_Wsin"]["sha256"]


def test_c2_misc1_4fm_pm(S, ref, tabs, gold):
    prg = ref.Program(scripts.C2_MISC1_4FM_PM)
    got = S.render(prg, srate=96000, tables=tabs)
    g = gold["config/C2_misc1_4fm_pm"]
    assert got.shape[0] == g["frames"] == 5760000
    assert gpuutil.sha(got) == g["sha256"]


@pytest.mark.parametrize("key,text", [
    ("config/C3_64v_1s", scripts.synth_c3(64, 1)),
    ("config/C3fm_64v_1s", scripts.synth_c3(64, 1, fm=True)),
    ("config/C4_48v_1s", scripts.synth_c4(48, 1)),
] + [(f"config/C5_script{i}", scripts.synth_c5_script(i)) for i in range(8)])
def test_synthetic_configs(S, ref, tabs, gold, key, text):
    prg = ref.Program(text)
    got = S.render(prg, srate=96000, tables=tabs)
    assert got.shape[0] == gold[key]["frames"]
    assert gpuutil.sha(got) == gold[key]["sha256"]
    # integer oscillator state at the end, bit-exact
    g = S.Generator(prg, 96000, tables=tabs)
    more = True
    while more:
        more, _, _ = g.run(24576)
    for op, want in enumerate(gold[key]["op_state"]):
        st = g.op_state(op)
        assert [st.inited, st.type, st.i0, st.i1, st.time] == want, (key, op)


@pytest.mark.parametrize("sched", [1, 2, 3])
def test_schedulers_agree_with_reference(S, ref, tabs, gold, sched):
    """One warp per voice (1), the ticketed persistent grid (2) and the balanced
    contiguous ranges (3, what the 4096-voice runs use) all reproduce the
    reference bit for bit."""
    for key, text in [("config/C3_64v_1s", scripts.synth_c3(64, 1)),
                      ("config/C3fm_64v_1s", scripts.synth_c3(64, 1, fm=True)),
                      ("config/C4_48v_1s", scripts.synth_c4(48, 1)),
                      ("config/C5_script3", scripts.synth_c5_script(3))]:
        prg = ref.Program(text)
        got = S.render(prg, srate=96000, tables=tabs, sched=sched)
        assert got.shape[0] == gold[key]["frames"], key
        assert gpuutil.sha(got) == gold[key]["sha256"], key
    feats = scripts.feature_scripts()
    for name in ["seq_update", "voices3", "regoal", "seq_overlap", "silence_mid", "pm_addrem"]:
        prg = ref.Program(feats[name])
        got = S.render(prg, srate=96000, tables=tabs, sched=sched, call_len=8192)
        g = gold.get("feat/" + name)
        assert g is None or gpuutil.sha(got) == g["sha256"], name
        assert np.array_equal(got, ref.render(prg, srate=96000, call_len=8192)), name


@pytest.mark.parametrize("call_len", [24576, 1024, 1000, 333, 77])
def test_state_after_every_call(S, ref, port, tabs, call_len):
    """All operator and voice state, bit for bit, after each call, any call size."""
    names = ["pm_chain", "fm_both", "self_w_mod", "self_r_pm", "seq_update", "voices3",
             "noise_am", "R_cub_self", "R_cub", "sweep_f_cub", "sweep_a_cub", "regoal", "pan_mod",
             "mod_finite", "self_w_off", "ratio_sweep", "noise_re", "noise_vi", "noise_bv",
             "wave_change", "pm_addrem", "seq_overlap", "silence_mid",
             "handover", "handover_pm", "handover_same_time", "handover_twice"]
    feats = scripts.feature_scripts()
    for name in names:
        prg = ref.Program(feats[name])
        gr = ref.RefGenerator(prg, 48000)
        gg = S.Generator(prg, 48000, tables=tabs, max_call_len=call_len)
        more, ncall = True, 0
        while more and ncall < 300:
            more, ba, na = gr.run(call_len)
            more2, bb, nb = gg.run(call_len)
            assert (more, na) == (more2, nb), (name, ncall)
            assert np.array_equal(ba, bb), (name, ncall)
            for op in range(prg.op_count):
                a = port.op_state_tuple(gr.op_state(op))
                b = port.op_state_tuple(gg.op_state(op))
                assert a == b, (name, ncall, op)
            for vo in range(prg.vo_count):
                assert gr.voice_state(vo)[:3] == gg.voice_state(vo)[:3], (name, ncall, vo)
            ncall += 1


def test_circular_modulator_graph(S, ref, port, tabs):
    """A program whose modulator lists loop back (sau/generator.c:685-690, ON_VISITED: the
    revisited operator renders zeros) -- built directly, the script language cannot state it."""
    prg = scripts.circular_program()
    for call_len in (24576, 1000):
        gr = ref.RefGenerator(prg, 96000)
        gg = S.Generator(prg, 96000, tables=tabs, max_call_len=call_len)
        more, ncall = True, 0
        while more:
            more, ba, na = gr.run(call_len)
            more2, bb, nb = gg.run(call_len)
            assert (more, na) == (more2, nb), ncall
            assert np.array_equal(ba, bb), ncall
            for op in range(prg.op_count):
                assert port.op_state_tuple(gr.op_state(op)) == port.op_state_tuple(gg.op_state(op)), (ncall, op)
            ncall += 1
        assert ncall >= 2


@pytest.mark.parametrize("sched", [1, 2, 3])
def test_operator_handover_all_schedulers(S, ref, tabs, sched):
    """An operator re-homed to another voice inside ONE call: the call is cut into separate
    render launches at the hand-over event (runtime.cpp:plan_call), under every scheduler and
    through the batched entry."""
    feats = scripts.feature_scripts()
    for name in ["handover", "handover_pm", "handover_same_time", "handover_twice"]:
        prg = ref.Program(feats[name])
        want = ref.render(prg, srate=96000)
        got = S.render(prg, srate=96000, tables=tabs, sched=sched)
        assert got.shape == want.shape and np.array_equal(got, want), (name, sched)
    if sched == 1:
        prgs = [ref.Program(feats[n]) for n in ["handover", "voices3", "handover_twice", "pm_chain"]]
        gens = [S.Generator(p, 96000, tables=tabs, max_call_len=24576) for p in prgs]
        refs = [ref.RefGenerator(p, 96000) for p in prgs]
        alive = [True] * len(gens)
        for _ in range(4):
            more, outs, lens = S.run_many(gens, 24576)
            for i, (g, r) in enumerate(zip(gens, refs)):
                if not alive[i]:
                    continue
                m, b, n = r.run(24576)
                assert (bool(more[i]), lens[i]) == (m, n), i
                assert np.array_equal(outs[i], b), i
                alive[i] = m
        for g in gens:
            g.close()


def test_balanced_scheduler_many_voices(S, ref, port, tabs):
    """More voices than resident warps (the shape auto-scheduling splits into
    balanced ranges with L2 hand-off of voice state between warps, in units of
    several blocks): 6000 voices with different durations, events inside the
    call, vs the oracle port."""
    import random
    rnd = random.Random(7)
    lines = ["S a.m0.003"]
    for i in range(6000):
        f = 110.0 * 2 ** rnd.uniform(0, 4)
        t = rnd.choice([0.02, 0.05, 0.11, 0.15, 0.3])
        lines.append(f"Wsin f{f:.3f} t{t} a1[g0.2 lxpe] c{rnd.uniform(-1, 1):.3f} "
                     f"p[Wtri r{rnd.choice([0.5, 1, 2])} a0.8[g0.1 llin]]")
    prg = ref.Program("\n".join(lines) + "\n")
    want = port.render(prg, srate=96000, tables=port.ref_tables())
    for sched in (0, 3):
        got = S.render(prg, srate=96000, tables=tabs, sched=sched)
        assert got.shape == want.shape
        assert np.array_equal(got, want), sched


def test_mono(S, ref, tabs, gold):
    prg = ref.Program(scripts.feature_scripts()["voices3"])
    got = S.render(prg, srate=44100, stereo=False, tables=tabs)
    assert got.shape[1] == 1
    assert gpuutil.sha(got) == gold["mono/voices3"]["sha256"]


def test_float_rows_within_1e5(S, ref, tabs):
    """Float carrier buffers vs the reference's gen_bufs[0] (block = call = 1024)."""
    for name in ["pm_chain", "fm_range", "self_w", "R_xpe_perlin", "am_range", "noise_gw"]:
        prg = ref.Program(scripts.feature_scripts()[name])
        gr = ref.RefGenerator(prg, 96000)
        gg = S.Generator(prg, 96000, tables=tabs, max_call_len=1024)
        for _ in range(6):
            _, _, n = gr.run(1024)
            gg.run(1024)
            want = gr.gen_buf(0)[:n] * np.float32(gr.amp_scale)
            s, _ = gg.voice_rows(0, n)
            tol = 1e-5 * np.maximum(np.abs(want), 1e-30)
            assert np.all(np.abs(s - want) <= tol), name
            assert np.array_equal(s.view(np.uint32), want.view(np.uint32)), name


# Modulator buffers (north_star: "float oscillator buffers within 1e-5").  The reference leaves, after a block, an
# operator's raw oscillator output (tmp_buf) and amplitude buffer in gen_bufs where nothing later reuses them
# (generator.c:548-605 buffer order); block_mix (generator.c:384-440) of those two is the operator's output.
# (op id, how the reference's output is rebuilt, gen_bufs indices)
MOD_TAPS = {
    "Wsin f200 t0.3 p[Wtri r2 a0.8[g0.1 llin] p[Wsqr r3.5 a0.5]]":
        [(0, "direct", 0), (1, "mul", 7, 6), (2, "mul", 10, 9)],
    "Wsin f300.r600[Wtri f7 a0.9] t0.3": [(0, "direct", 0), (1, "env", 8, 7)],
    "Wsin f300 a1.r0.2[Wsaw f6 p[Wtri f2 a0.3]] t0.3":
        [(0, "direct", 0), (1, "direct", 5), (1, "env", 9, 8), (2, "mul", 12, 11)],
    "Wsin f300[Wtri f11 a25] t0.3": [(0, "direct", 0), (1, "direct", 2)],
    "Wsin f200 t0.3 p.f[Wtri f50 a0.4]": [(0, "direct", 0), (1, "mul", 8, 7)],
    "Wsin f300 a1.r0.1[Rxpe f9 a0.7] t0.3": [(0, "direct", 0), (1, "direct", 5)],
    "Wcat f220 t0.5 p.a0.5[Wtri f1 a0.4]": [(0, "direct", 0), (1, "direct", 5)],
}


def test_modulator_buffers_within_1e5(S, ref, tabs):
    """Every operator's float output buffer (carrier AND modulators, through the debug tap of the general
    interpreter) vs the reference's gen_bufs, per 1024-frame block."""
    checked = 0
    for text, taps in MOD_TAPS.items():
        prg = ref.Program(text)
        gr = ref.RefGenerator(prg, 96000)
        gg = S.Generator(prg, 96000, tables=tabs, max_call_len=1024)
        gg.debug_tap()
        for call in range(6):
            _, pr, n = gr.run(1024)
            _, pg, ng = gg.run(1024)
            assert n == ng and np.array_equal(pr, pg), (text, call)
            for tap in taps:
                b = [gr.gen_buf(k)[:n] for k in tap[2:]]
                if tap[1] == "direct":
                    want = b[0]
                elif tap[1] == "mul":
                    want = b[0] * b[1]
                else:
                    s_amp = b[1] * np.float32(0.5)
                    want = b[0] * s_amp + np.abs(s_amp)
                got = gg.read_tap(tap[0], n)
                tol = 1e-5 * np.maximum(np.abs(want), 1e-30)
                assert np.all(np.abs(got - want) <= tol), (text, call, tap)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (text, call, tap)
                checked += 1
    assert checked == 6 * sum(len(t) for t in MOD_TAPS.values())


def test_run_many_matches_single(S, ref, tabs):
    """Batched entry point: same PCM as rendering each script alone."""
    texts = [scripts.synth_c5_script(i) for i in range(12)]
    prgs = [ref.Program(t) for t in texts]
    singles = [S.render(p, srate=96000, tables=tabs) for p in prgs]
    gens = [S.Generator(p, 96000, tables=tabs) for p in prgs]
    outs = [[] for _ in gens]
    alive = [True] * len(gens)
    while any(alive):
        more, pcm, lens = S.run_many(gens, 24576)
        for i in range(len(gens)):
            if alive[i]:
                outs[i].append(pcm[i][:2 * lens[i]].copy())
            alive[i] = alive[i] and bool(more[i])
    for i in range(len(gens)):
        got = np.concatenate(outs[i]).reshape(-1, 2)
        assert np.array_equal(got, singles[i]), i


def test_voice_sharded_mix_matches(S, ref, tabs):
    """Two voice shards + float-plane sum (what the NCCL reduce does across GPUs)
    == unsharded render within 1 LSB (summation order changes)."""
    import torch
    from saugns_b200.generator import planes_as_torch
    prg = ref.Program(scripts.synth_c3(32, 1))
    full = S.render(prg, srate=96000, tables=tabs)
    n = 24576
    ga = S.Generator(prg, 96000, tables=tabs, voice_range=(0, 16))
    gb = S.Generator(prg, 96000, tables=tabs, voice_range=(16, 32))
    out, more = [], True
    while more:
        ma, pa, la = ga.run_mix(n)
        mb, pb, lb = gb.run_mix(n)
        s = planes_as_torch(pa, 2 * n) + planes_as_torch(pb, 2 * n)
        torch.cuda.synchronize()
        pcm = ga.mix_to_pcm(s.data_ptr(), n)
        out.append(pcm[:2 * max(la, lb)])
        more = ma or mb
    got = np.concatenate(out).reshape(-1, 2)
    assert got.shape == full.shape
    assert np.abs(got.astype(np.int32) - full.astype(np.int32)).max() <= 1


def test_two_generators_one_program_alternating(S, ref, tabs):
    """Two generators alive on the same program at different rates, called alternately
    (Player_run with an audio device whose rate differs, saugns.c:585-599): instances
    must be fully independent."""
    prg = ref.Program(scripts.feature_scripts()["seq_overlap"])
    ra, rb = ref.RefGenerator(prg, 96000), ref.RefGenerator(prg, 44100)
    ga, gb = S.Generator(prg, 96000, tables=tabs), S.Generator(prg, 44100, tables=tabs)
    ma = mb = True
    while ma or mb:
        if ma:
            ma, want, n = ra.run(24576)
            m2, got, n2 = ga.run(24576)
            assert (ma, n) == (m2, n2) and np.array_equal(want, got)
        if mb:
            mb, want, n = rb.run(11289)
            m2, got, n2 = gb.run(11289)
            assert (mb, n) == (m2, n2) and np.array_equal(want, got)


def test_edge_calls(S, ref, tabs):
    """Zero-length calls, calls after the end of signal, out_len on the last call,
    mono/stereo switching between calls -- all as the reference answers them
    (generator.c:905-973)."""
    prg = ref.Program("Wsin f330 t0.07 Wtri f200 t0.03 c-0.5")
    gr, gg = ref.RefGenerator(prg, 48000), S.Generator(prg, 48000, tables=tabs, max_call_len=4096)
    for call_len, stereo in [(0, True), (1000, True), (0, False), (1000, False), (4096, True),
                             (4096, True), (333, True), (0, True)]:
        a = gr.run(call_len, stereo)
        b = gg.run(call_len, stereo)
        assert (a[0], a[2]) == (b[0], b[2]), (call_len, stereo)
        assert np.array_equal(a[1], b[1]), (call_len, stereo)
    # buf_len beyond the announced maximum is refused, not clipped
    with pytest.raises(RuntimeError):
        gg.run(4097)


def test_empty_and_silent_programs(S, ref, tabs):
    """A script with no sound at all, and one whose only voice has zero amplitude."""
    for text in ["S a.m0.5", "Wsin a0 t0.05", "Wsin t0"]:
        prg = ref.Program(text)
        want = ref.render(prg, srate=96000)
        got = S.render(prg, srate=96000, tables=tabs)
        assert got.shape == want.shape and np.array_equal(got, want), text


def test_many_short_voices_in_sequence(S, ref, port, tabs):
    """Voice slots reused by hundreds of events, several events inside one call (more
    inter-event segments per call than the initial tables hold: the growth path)."""
    import random
    rnd = random.Random(11)
    parts = []
    for i in range(300):
        parts.append(f"W{rnd.choice(scripts.WAVES)} f{rnd.uniform(100, 900):.2f} t0.004 "
                     f"a{rnd.uniform(0.1, 0.6):.2f} c{rnd.uniform(-1, 1):.2f} /0.002")
    prg = ref.Program("\n".join(parts) + "\n")
    want = port.render(prg, srate=96000, tables=port.ref_tables())
    got = S.render(prg, srate=96000, tables=tabs)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(want, ref.render(prg, srate=96000))


def test_nesting_limit_reported(S, ref, tabs):
    """Deeper operator nesting than the device interpreter holds fails at create with
    a message, it does not render garbage."""
    depth = 40
    text = "Wsin f200 t0.05 " + "".join("p[Wsin r1.01 a0.3 " for _ in range(depth)) + "]" * depth
    prg = ref.Program(text)
    with pytest.raises(RuntimeError, match="nesting too deep"):
        S.Generator(prg, 96000, tables=tabs)


def test_mix_stage_classes(S, ref, tabs):
    """The mix kernel's stage classes side by side (mix_kernel.cuh): 32-voice stages whose
    voices all run through a frame tile with constant pans (decision-free loop), stages with
    a voice ending inside the tile, with a moving pan, with fewer than 32 voices; stereo and
    mono, two call sizes.  Voice order of the float sums is the reference's: bit-exact."""
    import random
    rnd = random.Random(23)
    lines = []
    for i in range(107):                               # 3 full stages + one of 11 voices
        t = 0.30 if i < 64 else rnd.choice([0.30, 0.30, rnd.uniform(0.05, 0.29)])
        pan = f"c{rnd.uniform(-1, 1):.3f}"
        if i in (40, 70, 71, 100):
            pan = f"c-1[g1 t{rnd.uniform(0.05, 0.25):.3f}]"     # moving pan: r pieces
        lines.append(f"W{rnd.choice(scripts.WAVES)} f{rnd.uniform(80, 2000):.2f} t{t:.4f} "
                     f"a{rnd.uniform(0.2, 0.9):.2f} {pan}")
    prg = ref.Program("S a.m(1/107)\n" + "\n".join(lines) + "\n")
    for stereo in (True, False):
        want = ref.render(prg, srate=96000, stereo=stereo)
        for call_len in (24576, 1000):
            got = S.render(prg, srate=96000, tables=tabs, call_len=call_len, stereo=stereo)
            assert got.shape == want.shape and np.array_equal(got, want), (stereo, call_len)


def test_c3_full_voice_count_bit_exact(S, ref, tabs):
    """BASELINE config 3 at its full 4096 voices (one resident wave of 28-warp CTAs with
    the coefficient planes, the launch shape bench.py times), 0.6 s: every PCM sample and
    the integer state of every operator equal to the unmodified reference's."""
    prg = ref.Program(scripts.synth_c3(4096, 0.6, fm="mix"))
    want = ref.render(prg, srate=96000)
    g = S.Generator(prg, 96000, tables=tabs, max_call_len=24576)
    chunks, more = [], True
    while more:
        more, buf, n = g.run(24576)
        chunks.append(buf[:2 * n].copy())
    got = np.concatenate(chunks).reshape(-1, 2)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    gr = ref.RefGenerator(prg, 96000)
    more = True
    while more:
        more, _, _ = gr.run(24576)
    for op in range(0, prg.op_count, 37):
        a, b = gr.op_state(op), g.op_state(op)
        assert (a.i0, a.i1, a.time) == (b.i0, b.i1, b.time), op


def test_c4_full_voice_count_bit_exact(S, ref, tabs):
    """BASELINE config 4 at its full 1024 voices (self-PM W and R carriers, range-AM, ring
    modulation: the serial per-sample path), 0.3 s, bit-exact -- self-PM is chaotic, so
    this only holds if every feedback iteration reproduces the reference's float sequence."""
    prg = ref.Program(scripts.synth_c4(1024, 0.3))
    want = ref.render(prg, srate=96000)
    got = S.render(prg, srate=96000, tables=tabs)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_c3_full_size_longer(S, ref, tabs):
    """Config 3 at full size for 2 s (83 calls of 24 576 frames, past the end of every modulator's 1 s ramps
    and into the steady fused shapes bench.py times), streamed with run-ahead, bit-exact."""
    prg = ref.Program(scripts.synth_c3(4096, 2.0, fm="mix"))
    want = ref.render(prg, srate=96000)
    got = S.render(prg, srate=96000, tables=tabs)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_c4_full_size_longer(S, ref, tabs):
    """Config 4 at its full 1024 voices for 1.5 s, bit-exact."""
    prg = ref.Program(scripts.synth_c4(1024, 1.5))
    want = ref.render(prg, srate=96000)
    got = S.render(prg, srate=96000, tables=tabs)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_sweeps_past_2_24_samples(S, ref, tabs):
    """180 s sweeps at 96 kHz: line positions beyond 2**24 (SURVEY.md App. B.1's untested note), bit-exact."""
    prg = ref.Program(scripts.long_sweep_script())
    want = ref.render(prg, srate=96000)
    assert want.shape[0] > (1 << 24)
    got = S.render(prg, srate=96000, tables=tabs)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("knobs", [{"SAUGEN_MULTI": "0"}, {"SAUGEN_MULTI": "2"}, {"SAUGEN_MULTI": "3", "SAUGEN_TEAM": "5"},
                                   {"SAUGEN_MULTI": "8"}, {"SAUGEN_TEAM": "0"}])
def test_teams_over_several_ctas(S, knobs):
    """Few-voice scripts with nested FM under every way of spreading a voice: no teams, teams inside one CTA,
    teams over 2 / 3 (odd member counts) / up to 8 CTAs (render_team.cuh); each in its own process (the knobs
    are read once), every result bit-exact."""
    import subprocess
    import sys
    env = dict(os.environ, **knobs)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "gpu_multi_cta.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_kept_plans_equal_fresh_ones(S):
    """A voice's stable lowered plan is kept in global memory and reused call after call
    (render_kernel.cuh).  Under SAUGEN_PLAN_VERIFY=1 the kernel builds the plan afresh every time and
    compares it with the kept one wherever that would have been used: none may differ, across scripts
    with events, hand-overs, sweeps and two call sizes (a subprocess: the knob is read once)."""
    import subprocess
    import sys
    env = dict(os.environ, SAUGEN_PLAN_VERIFY="1")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "gpu_plan_verify.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_runahead_streaming_and_undo(S, ref, port, tabs):
    """A streaming caller (same call size again and again) gets the next call launched before it
    asks (runtime.cpp: run_call); a change of call size or channel count undoes the call that ran
    ahead, reading operator state in mid-stream sees the state of the last RETURNED call, and the
    end of the signal is reported by the right call."""
    feats = scripts.feature_scripts()
    for name in ["pm_chain", "seq_update", "voices3", "self_w_mod", "handover_twice", "fm_both", "regoal"]:
        prg = ref.Program(feats[name])
        gr = ref.RefGenerator(prg, 96000)
        gg = S.Generator(prg, 96000, tables=tabs, max_call_len=4096)
        sizes = [1024] * 5 + [777] * 3 + [1024] * 4 + [4096] * 100
        stereo = [True] * 6 + [False] * 4 + [True] * 200
        more, k = True, 0
        while more and k < len(sizes):
            more, ba, na = gr.run(sizes[k], stereo[k])
            more2, bb, nb = gg.run(sizes[k], stereo[k])
            assert (more, na) == (more2, nb), (name, k)
            assert np.array_equal(ba, bb), (name, k)
            if k in (3, 4, 9, 14):      # mid-stream inspection: the last returned call's state
                for op in range(prg.op_count):
                    assert port.op_state_tuple(gr.op_state(op)) == port.op_state_tuple(gg.op_state(op)), (name, k, op)
            k += 1
        assert not more, name
        for op in range(prg.op_count):   # final state: the call that ran past the end was undone
            assert port.op_state_tuple(gr.op_state(op)) == port.op_state_tuple(gg.op_state(op)), (name, op)
        more2, bb, nb = gg.run(1024)     # after the end: silence, no frames
        assert not more2 and nb == 0 and not bb.any()
        gg.close()


def test_runahead_device_calls_keep_two_buffers(S, ref, tabs):
    """run_device: the PCM of call k stays valid while call k+1 (already rendering) fills the other buffer."""
    import ctypes as C
    prg = ref.Program(scripts.synth_c3(48, 1, fm="mix"))
    want = ref.render(prg, srate=96000, call_len=8192)
    gg = S.Generator(prg, 96000, tables=tabs, max_call_len=8192)
    import torch
    got, more = [], True
    while more:
        more, ptr, n = gg.run_device(8192)
        t = torch.empty(8192 * 2, dtype=torch.int16, device="cuda")
        C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(ptr), C.c_size_t(8192 * 4), 3)
        got.append(t.cpu().numpy()[:2 * n])
    gg.close()
    assert np.array_equal(np.concatenate(got).reshape(-1, 2), want)


def test_time_split_teams_deep_fm(S, ref, port, tabs):
    """Few voices: steady stretches are split along time over a team of warps (render_team.cuh).
    Nested frequency modulation needs one counting pass per nesting level for the members' start
    phases; everything stays bit-exact, including the state after every call."""
    texts = {
        "fm3": "Wsin f300.r600[Wsin f7.r11[Wtri f0.5.r3[Wsin f0.31]]] t3 p[Wsin r2 a0.4]",
        "fm2_pm2": "Wsin f200.r320[Wsqr f3] t3 p[Wsin r1.5 a0.5 p[Wsin f90.r140[Wsin f2] a0.3] Wtri f50 a0.2]",
        "c2": scripts.C2_MISC1_4FM_PM.split("|")[0],
        "ratio_chain": "Wsin f110 t3 a1[g0.2 lxpe] p[Wtri r2 a0.8[g0.1 llin] p[Wsin r3.5 a0.5 p[Wsin r0.25 a0.3]]]",
        "five_voices": "\n".join(f"Wsin f{150 + 37 * i}.r{300 + 50 * i}[Wtri r0.{3 + i} a0.8[g0.1 llin]] t2.5 a0.5[g0.1 lxpe] "
                                 f"c{-0.8 + 0.4 * i} p[Wsin r{1 + i} a0.4]" for i in range(5)),
        "mixed_kinds": "Wsin f200 t2 p[Wsin r2 a0.5]\nNre t2 a0.2\nRlin f300 t2\nWsin f300 p.a0.6 t2\nWsaw f120[g400 lexp t1.5] t2",
        "slow_lfo": "Wsin f0.37 t3 a0.8\nWsin f440.r470[Wsin f0.11] t3",
    }
    for name, text in texts.items():
        prg = ref.Program(text)
        for call_len in (24576, 100000, 5000):
            want = ref.render(prg, srate=96000, call_len=call_len)
            got = S.render(prg, srate=96000, tables=tabs, call_len=call_len)
            assert got.shape == want.shape and np.array_equal(got, want), (name, call_len)
        gr = ref.RefGenerator(prg, 96000)
        gg = S.Generator(prg, 96000, tables=tabs, max_call_len=40000)
        more = True
        while more:
            more, ba, na = gr.run(40000)
            more2, bb, nb = gg.run(40000)
            assert (more, na) == (more2, nb) and np.array_equal(ba, bb), name
            for op in range(prg.op_count):
                assert port.op_state_tuple(gr.op_state(op)) == port.op_state_tuple(gg.op_state(op)), (name, op)
        gg.close()
