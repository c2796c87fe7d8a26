"""Built-in wave tables (saugns_b200/csrc/wavetab.cpp, strict IEEE) against the
tables the reference's own -ffast-math build produces (sau/wave.c)."""
import numpy as np


def test_builtin_tables_close_to_reference(ref):
    from saugns_b200 import generator
    tabs, coeffs = generator.builtin_tables()
    rt = ref.piluts()
    rc = ref.picoeffs()
    for w in range(12):
        assert abs(coeffs[w][0] - rc[w][0]) == 0 and abs(coeffs[w][1] - rc[w][1]) == 0
        assert coeffs[w][2] == rc[w][2]
        d = np.abs(tabs[w].astype(np.float64) - rt[w].astype(np.float64)).max()
        assert d <= 2.5e-7, (ref.WAVES[w], d)
    exact = [ref.WAVES[w] for w in range(12) if np.array_equal(tabs[w], rt[w])]
    # the tables that involve neither sqrtf nor a long running sum are identical
    assert "tri" in exact and "sqr" in exact
