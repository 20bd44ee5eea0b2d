"""Pins oracle/crossing.py (receiver-line crossing) to what the reference's own
Ray.addpoint returned (tests/golden/ft2_cross.npz)."""
import numpy as np

from _util import ft2_density


def cross_case(ft2_cross):
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(ft2_cross["nx"]), float(ft2_cross["n0"]))
    nf = Nfield(n, x=x, y=y)
    assert float(nf.st) == float(ft2_cross["st"])
    return nf


def test_crossing_oracle_matches_reference_addpoint(oracle_mod, ft2_cross):
    from oracle.crossing import crossing
    nf = cross_case(ft2_cross)
    la = ft2_cross["launch"]
    L = 4096
    tr = oracle_mod.trace_batch(la[:, 0].copy(), la[:, 1].copy(), la[:, 2].copy(), la[:, 3].copy(), L, nf.st,
                                nf.dx[0], nf.dy[0], nf.nx, nf.ny, nf.ax, nf.ay)
    assert np.array_equal(tr["nsteps"], ft2_cross["nkept"])
    res = ft2_cross["res"]
    seen = {"one": 0, "none": 0, "multi": 0, "extrap": 0}
    for li, (val, axis) in enumerate(zip(ft2_cross["line_val"], ft2_cross["line_axis"])):
        for r in range(la.shape[0]):
            n = int(tr["nsteps"][r])
            got = crossing(tr["traj"][r, :n], ft2_cross["x0"][r], ft2_cross["y0"][r], nf.dx[0], nf.dy[0], nf.st,
                           nf.ax, nf.ay, float(val), str(axis))
            want = res[li, r]
            assert got["count"] == int(want[0]), (li, r)
            if want[0] == 0:
                seen["none"] += 1
                continue
            seen["multi" if want[0] > 1 else "one"] += 1
            seen["extrap"] += got["extrapolated"]
            assert got["index"] == int(want[1])
            np.testing.assert_allclose([got[k] for k in ("x", "y", "kx", "ky", "lax", "ph")], want[2:8], rtol=1e-11, atol=1e-11)
    assert seen["one"] > 50 and seen["none"] > 20 and seen["multi"] >= 5 and seen["extrap"] >= 5, seen


def test_phase_lookup_oracle_matches_reference_interp(oracle_mod, ft2_cross):
    """oracle.crossing.phase_lookup == interp1d(r.ph, r.x, fill_value='extrapolate') evaluated on the
    reference's rays after their own addpoint (the neighbour step of Beam.field_r, beam.py:410-413)."""
    from oracle.crossing import phase_lookup
    nf = cross_case(ft2_cross)
    la = ft2_cross["launch"]
    tr = oracle_mod.trace_batch(la[:, 0].copy(), la[:, 1].copy(), la[:, 2].copy(), la[:, 3].copy(), 4096, nf.st,
                                nf.dx[0], nf.dy[0], nf.nx, nf.ny, nf.ax, nf.ay)
    checked = 0
    for li, (val, axis) in enumerate(zip(ft2_cross["line_val"], ft2_cross["line_axis"])):
        for r in range(la.shape[0]):
            t = ft2_cross["look_t"][li, r]
            if np.isnan(t[0]):
                continue
            n = int(tr["nsteps"][r])
            got = phase_lookup(tr["traj"][r, :n], ft2_cross["x0"][r], ft2_cross["y0"][r], nf.dx[0], nf.dy[0], nf.st,
                               nf.ax, nf.ay, float(val), str(axis), t)
            np.testing.assert_allclose(np.ravel(got), ft2_cross["look_xy"][li, r], rtol=1e-10, atol=1e-10)
            checked += 1
    assert checked > 60
