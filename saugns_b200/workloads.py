"""Synthetic many-voice workloads (BASELINE.json configs[2]: "4096 concurrent
voices of 3-operator PM/FM chains with envelope ramps, 60 s at 96 kHz").

`synth_c3` writes the workload as a SAU script (for the reference front end and
the CPU baseline); `build_c3` emits the identical sauProgram directly
(saugns_b200.program.ProgramBuilder) so that the GPU arm of bench.py needs no
script front end.  tests/test_program_builder.py checks the two agree field by
field.  fm: False = PM chain, True = range-FM + PM, "mix" = alternating.
"""
import random

WAVES = ["sin", "tri", "srs", "sqr", "ean", "cat", "eto", "par", "mto", "saw", "hsi", "spa"]
LINES = ["cos", "lin", "sah", "exp", "log", "xpe", "lge", "sqe", "cub", "smo", "ncl", "nhl", "uwh"]
NOISES = ["wh", "gw", "bw", "tw", "re", "vi", "bv"]


def _is_fm(fm, i):
    return (i % 2 == 1) if fm == "mix" else bool(fm)


def synth_c3(n_voices=4096, secs=60, seed=1, fm=False):
    """BASELINE config 3: n voices of 3-operator PM (or FM) chains with ramps."""
    rnd = random.Random(seed)
    lines = [f"S a.m{0.3 / n_voices ** 0.5:.6f}"]
    for i in range(n_voices):
        f = 110.0 * 2 ** rnd.uniform(0, 4)
        c = rnd.uniform(-1, 1)
        r1 = rnd.choice([0.5, 1, 1.5, 2, 3])
        r2 = rnd.choice([1, 2, 3.5, 7])
        if _is_fm(fm, i):
            lines.append(
                f"Wsin f{f:.3f}.r{2 * f:.3f}[Wtri r{r1} a0.8[g0.1 llin]] t{secs} "
                f"a1[g0.2 lxpe] c{c:.3f} p[Wsin r{r2} a0.5]")
        else:
            lines.append(
                f"Wsin f{f:.3f} t{secs} a1[g0.2 lxpe] c{c:.3f} "
                f"p[Wtri r{r1} a0.8[g0.1 llin] p[Wsin r{r2} a0.5]]")
    return "\n".join(lines) + "\n"



def build_c3(n_voices=4096, secs=60, seed=1, fm=False):
    """The same program synth_c3() describes, built without a script front end
    (saugns_b200.program.ProgramBuilder); used by bench.py's product arm."""
    from saugns_b200 import program as P
    rnd = random.Random(seed)
    pb = P.ProgramBuilder(ampmult=float(f"{0.3 / n_voices ** 0.5:.6f}"))
    ms = int(round(secs * 1000))
    for i in range(n_voices):
        f_raw = 110.0 * 2 ** rnd.uniform(0, 4)
        f = float(f"{f_raw:.3f}")
        c = float(f"{rnd.uniform(-1, 1):.3f}")
        r1 = rnd.choice([0.5, 1, 1.5, 2, 3])
        r2 = rnd.choice([1, 2, 3.5, 7])
        m2 = P.ProgramBuilder.wave("sin", freq=P.value(r2, ratio=True), amp=0.5)
        if _is_fm(fm, i):
            m1 = P.ProgramBuilder.wave("tri", freq=P.value(r1, ratio=True),
                                       amp=P.value(0.8, goal=0.1, line="lin"))
            carr = P.ProgramBuilder.wave(
                "sin", freq=f, freq2=float(f"{2 * f_raw:.3f}"), time_ms=ms, pan=c,
                amp=P.value(1.0, goal=0.2, line="xpe"), mods={"rfmod": [m1], "pmod": [m2]})
        else:
            m1 = P.ProgramBuilder.wave("tri", freq=P.value(r1, ratio=True),
                                       amp=P.value(0.8, goal=0.1, line="lin"), mods={"pmod": [m2]})
            carr = P.ProgramBuilder.wave("sin", freq=f, time_ms=ms, pan=c,
                                         amp=P.value(1.0, goal=0.2, line="xpe"), mods={"pmod": [m1]})
        pb.add_voice(carr)
    return pb.finish()


def synth_c4(n_voices=1024, secs=60, seed=2):
    """BASELINE config 4: self-feedback PM carriers with range-AM / ring-mod."""
    rnd = random.Random(seed)
    lines = [f"S a.m{0.3 / n_voices ** 0.5:.6f}"]
    for i in range(n_voices):
        f = 110.0 * 2 ** rnd.uniform(0, 4)
        c = rnd.uniform(-1, 1)
        pa = rnd.uniform(0.3, 1.0)
        fm = rnd.uniform(0.5, 8)
        k = i % 3
        if k == 0:
            lines.append(f"Wsin f{f:.3f} t{secs} p.a{pa:.3f} a0.5.r1[Wsin f{fm:.3f}] c{c:.3f}")
        elif k == 1:
            lines.append(f"Rlin f{f:.3f} t{secs} p.a{pa:.3f} a0.5.r1[Wsin f{fm:.3f}] c{c:.3f}")
        else:
            lines.append(f"Wtri f{f:.3f} t{secs} p.a{pa:.3f} a0[Wsin f{fm * 20:.3f} a0.8] c{c:.3f}")
    return "\n".join(lines) + "\n"


def build_c4(n_voices=1024, secs=60, seed=2):
    """The program synth_c4() describes, built without a script front end."""
    from saugns_b200 import program as P
    B = P.ProgramBuilder
    rnd = random.Random(seed)
    pb = B(ampmult=float(f"{0.3 / n_voices ** 0.5:.6f}"))
    ms = int(round(secs * 1000))
    for i in range(n_voices):
        f = float(f"{110.0 * 2 ** rnd.uniform(0, 4):.3f}")
        c = float(f"{rnd.uniform(-1, 1):.3f}")
        pa = float(f"{rnd.uniform(0.3, 1.0):.3f}")
        fm = rnd.uniform(0.5, 8)
        k = i % 3
        if k == 0:
            carr = B.wave("sin", freq=f, time_ms=ms, pm_a=pa, amp=0.5, amp2=1.0, pan=c,
                          mods={"ramod": [B.wave("sin", freq=float(f"{fm:.3f}"))]})
        elif k == 1:
            carr = B.raseg("lin", freq=f, time_ms=ms, pm_a=pa, amp=0.5, amp2=1.0, pan=c,
                           mods={"ramod": [B.wave("sin", freq=float(f"{fm:.3f}"))]})
        else:
            carr = B.wave("tri", freq=f, time_ms=ms, pm_a=pa, amp=0.0, pan=c,
                          mods={"amod": [B.wave("sin", freq=float(f"{fm * 20:.3f}"), amp=0.8)]})
        pb.add_voice(carr)
    return pb.finish()


def build_c5_script(index):
    """The program synth_c5_script(index) describes, built without a script front end."""
    from saugns_b200 import program as P
    B = P.ProgramBuilder
    rnd = random.Random(1000 + index)
    nv = rnd.randint(4, 16)
    pb = B(ampmult=float(f"{0.3 / nv ** 0.5:.6f}"))
    for _ in range(nv):
        t = rnd.uniform(1, 10)
        ms = int(round(float(f"{t:.3f}") * 1000))
        f_raw = 110.0 * 2 ** rnd.uniform(0, 4)
        f = float(f"{f_raw:.3f}")
        c = float(f"{rnd.uniform(-1, 1):.3f}")
        kind = rnd.randrange(4)
        if kind == 0:
            w, w2 = rnd.choice(WAVES), rnd.choice(WAVES)
            r = rnd.choice([0.5, 1, 2, 3])
            a = float(f"{rnd.uniform(0.1, 1):.3f}")
            carr = B.wave(w, freq=f, time_ms=ms, pan=c,
                          mods={"pmod": [B.wave(w2, freq=P.value(r, ratio=True), amp=a)]})
        elif kind == 1:
            n = rnd.choice(NOISES)
            carr = B.noise(n, time_ms=ms, pan=c, amp=float(f"{rnd.uniform(0.1, 0.8):.3f}"))
        elif kind == 2:
            mode = rnd.choice("ugbtfa") + rnd.choice(["", "h", "p", "s", "v", "z"])
            carr = B.raseg(rnd.choice(LINES), mode, freq=f, time_ms=ms, pan=c)
        else:
            w = rnd.choice(WAVES)
            g = float(f"{f_raw * rnd.uniform(0.5, 2):.3f}")
            ln = rnd.choice(LINES)
            carr = B.wave(w, freq=P.value(f, goal=g, line=ln), time_ms=ms, pan=c, amp=1.0, amp2=0.0,
                          mods={"ramod": [B.wave("sin", freq=float(f"{rnd.uniform(0.5, 9):.3f}"))]})
        pb.add_voice(carr)
    return pb.finish()


def synth_c5_script(index):
    """BASELINE config 5: one of the independent mixed scripts (seed 1000+index)."""
    rnd = random.Random(1000 + index)
    nv = rnd.randint(4, 16)
    lines = [f"S a.m{0.3 / nv ** 0.5:.6f}"]
    for _ in range(nv):
        t = rnd.uniform(1, 10)
        f = 110.0 * 2 ** rnd.uniform(0, 4)
        c = rnd.uniform(-1, 1)
        kind = rnd.randrange(4)
        if kind == 0:
            w, w2 = rnd.choice(WAVES), rnd.choice(WAVES)
            lines.append(f"W{w} f{f:.3f} t{t:.3f} c{c:.3f} p[W{w2} r{rnd.choice([0.5, 1, 2, 3])} "
                         f"a{rnd.uniform(0.1, 1):.3f}]")
        elif kind == 1:
            lines.append(f"N{rnd.choice(NOISES)} t{t:.3f} c{c:.3f} a{rnd.uniform(0.1, 0.8):.3f}")
        elif kind == 2:
            mode = rnd.choice("ugbtfa") + rnd.choice(["", "h", "p", "s", "v", "z"])
            lines.append(f"R{rnd.choice(LINES)} m{mode} f{f:.3f} t{t:.3f} c{c:.3f}")
        else:
            lines.append(f"W{rnd.choice(WAVES)} f{f:.3f}[g{f * rnd.uniform(0.5, 2):.3f} "
                         f"l{rnd.choice(LINES)}] t{t:.3f} c{c:.3f} a1.r0[Wsin f{rnd.uniform(0.5, 9):.3f}]")
    return "\n".join(lines) + "\n"


C2_TEXT = """Wsin t15 f500.r501[Wsin f1] p[
	Wsin f400.r800[
		Wsqr f1.r10[Wsin f5000]
		Wtri f0.1.r10.0[Wsin f0.2]
	]
] |

Wsin t15 f400.r500[Wsqr f10] p[
	Wsin r1.22/2 a.5
	Wsin f244 a.5
] |

Wsin t15 f600.r666[Wsin f2] p[
	Wsin f400
	Wsin f400.r500[Wsin f.1]
] |

Wsin t15 f222.r666[Wsin f0.1] p[
	Wsin r2/1
	Wsin r4/3
	Wsin r3/7
]
"""


def build_c2():
    """BASELINE config 2 (the reference's examples/misc1-4fm_pm.sau, quoted above): four 15 s voices one after
    the other in ONE voice slot, range-FM nested up to three deep under a PM modulator -- built without a script
    front end (bench.py's product arm); the same program the reference's parser makes of the text."""
    from saugns_b200 import program as P
    B = P.ProgramBuilder
    W = B.wave
    r = lambda x: P.value(x, ratio=True)
    pb = B(ampmult=1.0, amp_div_voices=True)
    pb.add_voice(W("sin", time_ms=15000, freq=500.0, freq2=501.0, mods={
        "rfmod": [W("sin", freq=1.0)],
        "pmod": [W("sin", freq=400.0, freq2=800.0, mods={"rfmod": [
            W("sqr", freq=1.0, freq2=10.0, mods={"rfmod": [W("sin", freq=5000.0)]}),
            W("tri", freq=0.1, freq2=10.0, mods={"rfmod": [W("sin", freq=0.2)]})]})]}), vo_id=0)
    pb.add_voice(W("sin", time_ms=15000, freq=400.0, freq2=500.0, mods={
        "rfmod": [W("sqr", freq=10.0)],
        "pmod": [W("sin", freq=r(1.22 / 2), amp=0.5), W("sin", freq=244.0, amp=0.5)]}), wait_ms=15000, vo_id=0)
    pb.add_voice(W("sin", time_ms=15000, freq=600.0, freq2=666.0, mods={
        "rfmod": [W("sin", freq=2.0)],
        "pmod": [W("sin", freq=400.0),
                 W("sin", freq=400.0, freq2=500.0, mods={"rfmod": [W("sin", freq=0.1)]})]}), wait_ms=15000, vo_id=0)
    pb.add_voice(W("sin", time_ms=15000, freq=222.0, freq2=666.0, mods={
        "rfmod": [W("sin", freq=0.1)],
        "pmod": [W("sin", freq=r(2 / 1)), W("sin", freq=r(4 / 3)), W("sin", freq=r(3 / 7))]}), wait_ms=15000, vo_id=0)
    return pb.finish()
