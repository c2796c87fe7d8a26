"""CPU checks of the boundary: the C-ABI library loads, exports every symbol
include/saugen_b200.h declares, and its mirror of the reference's program
data model has the reference's exact layout."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_all_declared_symbols():
    import saugns_b200
    L = saugns_b200.lib()
    decl = open(os.path.join(ROOT, "include", "saugen_b200.h")).read()
    decl += open(os.path.join(ROOT, "include", "sau_program_abi.h")).read()
    names = set(re.findall(r"\b(saugen_[a-z_]+)\s*\(", decl))
    assert len(names) >= 14
    for n in sorted(names):
        assert hasattr(L, n), n


def test_abi_layout_matches_reference(ref):
    from saugns_b200 import generator
    assert generator.abi_layout() == ref.abi_layout()


def test_dropin_object_exports_reference_symbols():
    import subprocess
    obj = os.path.join(ROOT, "saugns_b200", "dropin.o")
    out = subprocess.check_output(["nm", obj], text=True)
    for sym in ["sau_create_Generator", "sau_destroy_Generator", "sauGenerator_run", "sauNoise_names"]:
        assert re.search(r"\b[TDR] " + sym + r"\b", out), sym


def test_no_device_fails_loudly():
    import saugns_b200
    if saugns_b200.device_count() > 0:
        return
    class P:  # noqa: E306
        ptr = 1
    try:
        saugns_b200.Generator(P(), 96000)
    except RuntimeError as e:
        assert "no usable CUDA device" in str(e)
    else:
        raise AssertionError("creation must fail without a GPU")
