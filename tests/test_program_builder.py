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
