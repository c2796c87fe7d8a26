"""Script corpora shared by the CPU (oracle) and GPU (parity) tests.

FEATURE_SCRIPTS are short single-feature SAU scripts that together touch every
arithmetic path of the generator back end: the 12 wave
types, 7 noise types, 6 R functions x flags, 13 line shapes as parameter
sweeps and as R segment shapes, PM / fPM / FM / range-FM / AM / range-AM,
self-PM for W and R, panning (constant, swept, modulated), the A operator,
finite-time modulators, event sequencing and voice reuse.
"""
import random

from saugns_b200.workloads import (synth_c3, build_c3, synth_c4, synth_c5_script,  # noqa: F401
                                   build_c4, build_c5_script,
                                   WAVES, LINES, NOISES)   # the BASELINE workloads live with the product



def feature_scripts():
    s = {}
    for w in WAVES:
        s[f"wave_{w}"] = f"W{w} f220 t0.25"
        s[f"wave_{w}_pm"] = f"W{w} f330 t0.2 p[Wsin f440 a0.7]"
    for n in NOISES:
        s[f"noise_{n}"] = f"N{n} t0.2 a0.5"
    for l in LINES:
        s[f"sweep_f_{l}"] = f"Wsin f200[g900 l{l} t0.21] t0.3"
        s[f"sweep_a_{l}"] = f"Wtri f300 a1[g0.1 l{l} t0.2] t0.3"
        s[f"R_{l}"] = f"R{l} f300 t0.2"
        s[f"R_{l}_perlin"] = f"R{l} mp f300 t0.2"
        s[f"R_{l}_self"] = f"R{l} f250 p.a0.6 t0.2"
    for m in ["u", "g", "b", "b3", "t", "t4", "f", "f2", "f5", "a", "uv", "bv", "b2v", "fv", "f3v",
              "uh", "gz", "us", "uhp", "b4hz", "t2ps", "f1vs", "gp", "ahz", "up", "uvp", "uzp"]:
        s[f"Rmode_{m}"] = f"Rlin m{m} f260 t0.15"
        s[f"Rmode_{m}_self"] = f"Rsqe m{m} f260 p.a0.5 t0.15"
    s["Rmode_a_alpha"] = "Rlin ma.a0.7548776662 f300 t0.2"
    s["pm_chain"] = "Wsin f200 t0.3 p[Wtri r2 a0.8[g0.1 llin] p[Wsin r3.5 a0.5]]"
    s["pm_two"] = "Wsin f200 t0.3 p[Wsin f300 a0.5 Wsaw f120 a0.2]"
    s["fpm"] = "Wsin f300 t0.3 p.f[Wsin f100 a0.9]"
    s["pm_fpm"] = "Wsin f300 t0.3 p[Wsin f150 a0.4] p.f[Wtri f90 a0.5]"
    s["fm_add"] = "Wsin f300[Wsin f5 a40] t0.3"
    s["fm_range"] = "Wsin f300.r600[Wsin f7] t0.3"
    s["fm_range2"] = "Wsin f300.r600[Wsin f7 Wtri f3 a-0.8] t0.3"
    s["fm_both"] = "Wsin f300.r500[Wsin f4][Wsin f11 a25] t0.3"
    s["am_add"] = "Wsin f300 a0.5[Wsin f6 a0.4] t0.3"
    s["am_range"] = "Wsin f300 a1.r0[Wsin f6] t0.3"
    s["rm"] = "Wsin f300 a0[Wsin f80] t0.3"
    s["am_range_sweep"] = "Wsin f300 a1.r0.2[g0.9 lsqe t0.2][Wsin f6] t0.3"
    s["ratio_sweep"] = "Wsin f200 t0.3 p[Wsin r2[g3 lexp t0.25] a0.6]"
    s["ratio_fm"] = "Wsin f200 t0.3 f[Wsin r0.5 a30]"
    s["self_w"] = "Wsin f220 p.a0.8 t0.3"
    s["self_w_sweep"] = "Wsin f220 p.a0[g1.2 t0.2] t0.3"
    s["self_w_mod"] = "Wcat f220 t0.5 p.a0.5[Wsin f1 a0.4]"
    s["self_w_off"] = "Wsin f220 p.a0.7[g0 t0.05] t0.3"
    s["self_r"] = "Rlin f220 p.a0.7 t0.3"
    s["self_r_pm"] = "Rcos mg f220 p.a0.7 p[Wsin f100 a0.3] t0.3"
    s["r_pm_fm"] = "Rsmo f200.r400[Wsin f3] p[Wsin f300 a0.5] t0.3"
    s["r_fpm"] = "Rlin mh f200 p.f[Wsin f50 a0.7] t0.3"
    s["pan_const"] = "Wsin f300 c-0.5 t0.2"
    s["pan_sweep"] = "Wsin f300 c-1[g1 t0.15] t0.3"
    s["pan_mod"] = "Wsin f300 c0[Wsin f5 a0.8] t0.3"
    s["pan_mod_ratio"] = "Wsin f300 c0.1[Wsin r0.01 a0.8] t0.3"
    s["pan_mod_ratio_self"] = "Wsin f300 p.a0.5 c0.1[Wsin r0.01 a0.8] t0.3"
    s["pan_mod_ratio_pm"] = "Wsin f300 c0.1[Wsin r0.01 a0.8] p[Wsin f200 a0.3] t0.3"
    s["pan_mod_ratio_fm"] = "Wsin f300.r400[Wsin f3] c0.1[Wtri r0.02 a0.6] t0.3"
    s["amp_op"] = "A0[Wsin f300 a0.5 Wtri f100 a0.3] t0.3"
    s["amp_op_range"] = "A0.8.r0.1[Wsin f8] t0.2"
    s["noise_am"] = "Nwh a0.5.r0[Wsin f9] t0.3"
    s["noise_as_pm"] = "Wsin f300 p[Nre a0.2] t0.3"
    s["mod_finite"] = "Wsin f200 t0.5 p[Wsin f300 t0.1 a0.8]"
    s["mod_finite_fm"] = "Wsin f200[Wsin f9 a50 t0.2] t0.5"
    s["voices3"] = "Wsin f200 c-0.3 t0.3 Wtri f301 c0.4 t0.2 Nwh a0.2 t0.25"
    s["seq_update"] = "Wsin f200 t0.3; f300 t0.2; f150[g400] t0.25"
    s["seq_bar"] = "Wsin f200 t0.2 | Wtri f300 t0.15 | Nvi t0.1"
    s["seq_delay"] = "Wsin f200 t0.2 /0.3 Wsaw f100 t0.2"
    s["seq_overlap"] = "Wsin f200 t0.4 /0.1 Wsin f250 t0.4 /0.1 Wsin f300 t0.4"
    s["regoal"] = "Wsin f200[g800 t0.4] t0.2; f[g100 t0.1] t0.2"
    s["wave_change"] = "Wsin f200 t0.2; wsaw t0.2; wtri p0.25 t0.1"
    s["pm_addrem"] = "Wsin f200 t0.2 p[Wsin f300]; p-[] t0.1; p[Wtri f100 a0.5] t0.2"
    s["ampmult"] = "S a0.3 Wsin f200 t0.2 Wsin f300 t0.2"
    s["zero_freq"] = "Wsin f0 t0.1 p[Wsin f200 a0.5]"
    s["neg_freq"] = "Wsaw f-200 t0.2"
    s["high_freq"] = "Wsqr f30000 t0.1"
    s["deep"] = "Wsin f200 t0.3 p[Wsin r2 p[Wsin r2 p[Wsin r2 p[Wsin r0.5 a0.3]]]]"
    s["silence_mid"] = "Wsin f200 t0.1 /0.6 Wsin f300 t0.1"
    # an operator handed over between voices: its voice slot is recycled by another carrier, a later
    # `@label` update then gives the same op id a NEW voice (parseconv); both events in one call
    s["handover"] = "'a Wsin f440 t0.01\n/0.02 Wsin f220 t0.5\n/0.01 @a t0.5 f880"
    s["handover_pm"] = ("'a Wsin f300 t0.02 p[Wtri f70 a0.6]\n/0.03 Wsaw f110 t0.3 a0.5\n"
                        "/0.001 @a t0.3 f450[g200 lexp t0.2]\n/0.1 @a a0.3 t0.1")
    s["handover_same_time"] = "'a Wsin f440 t0.01\n/0.02 Wsin f220 t0.5 @a t0.4 f660"
    s["handover_twice"] = ("'a Wsin f440 t0.01 'b Wtri f330 t0.01\n/0.02 Wsin f220 t0.3 Wsin f275 t0.3\n"
                           "/0.01 @b t0.3 f700\n/0.005 @a t0.3 f880\n/0.2 @b f350 t0.1")
    return s


def circular_program():
    """A modulator graph with a circular reference (carrier <- tri <- sin <- carrier), which the
    script language cannot state but the generator guards against (ON_VISITED, sau/generator.c:
    685-690: the revisited operator renders zeros); built directly as a sauProgram."""
    from saugns_b200 import program as P
    pb = P.ProgramBuilder(ampmult=1.0)
    m2 = P.ProgramBuilder.wave("sin", freq=50.0, amp=0.4, raw_mods={"pmod": [0]})   # back to the carrier
    m1 = P.ProgramBuilder.wave("tri", freq=300.0, amp=0.7, mods={"pmod": [m2]})
    pb.add_voice(P.ProgramBuilder.wave("sin", freq=200.0, time_ms=300, pan=0.2, mods={"pmod": [m1]}))
    # a second voice whose only modulator is itself
    pb.add_voice(P.ProgramBuilder.wave("saw", freq=120.0, time_ms=200, pan=-0.4, raw_mods={"pmod": [3], "fmod": [3]}))
    return pb.finish()


# BASELINE config 2: the reference's examples/misc1-4fm_pm.sau (saugns v0.4.7,
# by Joel K. Pettersson; input data for the parity test, quoted verbatim so
# that the GPU box, which has no /root/reference, can render it).
C2_MISC1_4FM_PM = """Wsin t15 f500.r501[Wsin f1] p[
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


def long_sweep_script():
    """Sweeps that run past 2**24 samples at 96 kHz (the as-compiled vector body converts the
    unsigned position in two halves there), on the shapes that use the unsigned and the signed position."""
    return ("Wsin f100[g1000 lxpe t180] t180 a0.5 c-0.5\n"
            "Wtri f2000[g50 llge t178] t180 a0.3[g0.05 lsmo t179] c0.5\n"
            "Wsaw f300[g600 lcub t177] t180 a0.2[g0.1 lsqe t179.5]\n")
