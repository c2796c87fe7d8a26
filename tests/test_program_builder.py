"""ProgramBuilder must emit the same sauProgram event/operator data as the
reference front end does for the equivalent script (bench.py's product arm
builds its workload with it instead of parsing)."""
import scripts


def _strip(d):
    return d


def test_builder_matches_parser_c3(ref):
    from saugns_b200 import program as P
    for fm in (False, True, "mix"):
        prg = ref.Program(scripts.synth_c3(24, 2, seed=5, fm=fm))
        built = scripts.build_c3(24, 2, seed=5, fm=fm)
        a = P.dump(prg.ptr)
        b = P.dump(built.ptr)
        assert a["events"] == b["events"]
        for k in ("mode", "vo_count", "op_count", "op_nest_depth", "duration_ms", "ampmult"):
            assert a[k] == b[k], k


def test_builder_program_renders_identically_on_oracle(ref, port):
    import numpy as np
    prg = ref.Program(scripts.synth_c3(6, 1, seed=3))
    built = scripts.build_c3(6, 1, seed=3)
    a = port.render(prg, srate=48000)
    b = port.render(built, srate=48000)
    assert np.array_equal(a, b)


def _same_program(P, a_ptr, b_ptr):
    a, b = P.dump(a_ptr), P.dump(b_ptr)
    assert a["events"] == b["events"]
    for k in ("mode", "vo_count", "op_count", "op_nest_depth", "duration_ms", "ampmult"):
        assert a[k] == b[k], k


def test_builder_matches_parser_c4(ref):
    """BASELINE config 4 (self-PM W and R carriers, range-AM, ring modulation): bench.py's C4 leg."""
    from saugns_b200 import program as P
    prg = ref.Program(scripts.synth_c4(96, 60))
    built = scripts.build_c4(96, 60)
    _same_program(P, prg.ptr, built.ptr)


def test_builder_matches_parser_c5(ref):
    """BASELINE config 5 (mixed W+PM / N / R modes / swept W with range-AM scripts, seeded operators
    drawing from the front end's deterministic SplitMix32): bench.py's C5 leg."""
    from saugns_b200 import program as P
    for i in list(range(0, 200)) + list(range(1250, 10000, 173)):
        prg = ref.Program(scripts.synth_c5_script(i))
        built = scripts.build_c5_script(i)
        _same_program(P, prg.ptr, built.ptr)


def test_builder_programs_render_identically_on_oracle_c4_c5(ref, port):
    import numpy as np
    for prg, built in [(ref.Program(scripts.synth_c4(12, 1)), scripts.build_c4(12, 1)),
                       (ref.Program(scripts.synth_c5_script(7)), scripts.build_c5_script(7))]:
        a = port.render(prg, srate=48000, max_frames=48000)
        b = port.render(built, srate=48000, max_frames=48000)
        assert np.array_equal(a, b)


def test_builder_c2_equals_the_parsed_script(ref, port):
    """BASELINE config 2 (examples/misc1-4fm_pm.sau): four voices one after the other in ONE voice slot --
    the builder's program is the parser's, field for field, and renders the same PCM."""
    import json
    import numpy as np
    from saugns_b200 import workloads, program as P
    import scripts
    assert workloads.C2_TEXT.split() == scripts.C2_MISC1_4FM_PM.split()
    built, parsed = workloads.build_c2(), ref.Program(workloads.C2_TEXT)
    db, dp = P.dump(built.ptr), P.dump(parsed.ptr)
    assert json.dumps(db, sort_keys=True, default=str) == json.dumps(dp, sort_keys=True, default=str)
    assert np.array_equal(port.render(built, srate=96000), ref.render(parsed, srate=96000))
