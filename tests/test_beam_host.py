"""Host-side mirror of the reference's Beam / Ray.addpoint against fixtures
produced by the unmodified beam.py / ray.py (tests/golden/beam.npz, ft2_cross.npz).  No GPU."""
import numpy as np
import pytest
from scipy.interpolate import interp1d

from _util import ft2_density
from rayrtrace_b200.nfield import Nfield


def _beam(beam_fix, **kw):
    from rayrtrace_b200.beam import Beam
    return Beam(x0=float(beam_fix["x0"]), y0=float(beam_fix["y0"]), kx0=0, ky0=1, **kw)


def test_analytic_gaussian_beam_matches_reference(beam_fix):
    b = _beam(beam_fix)
    assert b.wave_len == float(beam_fix["wave_len"])
    assert b.rayleigh_len(n=0.97) == float(beam_fix["rayleigh"])
    fs = b.field(x=beam_fix["probe_x"], y=beam_fix["probe_y"], n=0.97, outfmt="dict")
    assert fs["amp"].tobytes() == beam_fix["probe_amp"].tobytes()
    assert fs["phase"].tobytes() == beam_fix["probe_phase"].tobytes()
    assert fs["field"].tobytes() == beam_fix["probe_field"].tobytes()
    xb, yb = b.lab2beam(1.0, -7.0)
    xl, yl = b.beam2lab(xb, yb)
    assert (xl, yl) == pytest.approx((1.0, -7.0), abs=1e-15)


def test_wavefront_matches_the_reference_finder(beam_fix):
    """The reference finds the front with skimage.measure.find_contours on a 1000 x 1000 phase map
    (beam.py:225-247); the fixture holds the vertices its own code produced (on the restated
    find_contours of oracle/refenv.py).  The product's default finder must return the same vertices
    from the narrow band of map nodes it evaluates, and from the full map."""
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    b = _beam(beam_fix)
    assert b.front_method == "contour"
    wf = b.wavefront_r(dist=3, ampbnd=1e-3, n_field=nf)
    assert wf.x.tobytes() == beam_fix["wf_x"].tobytes()
    assert wf.y.tobytes() == beam_fix["wf_y"].tobytes()
    rng = np.random.default_rng(0)
    for _ in range(4):       # band == full map, for other beams too
        from rayrtrace_b200.beam import Beam
        bb = Beam(x0=rng.uniform(-4, 4), y0=-9.0, kx0=0, ky0=1, waist=rng.uniform(0.25, 0.6), focus=rng.uniform(-2, 0.5),
                  freq=rng.uniform(100e9, 160e9))
        dist, ne = rng.uniform(1.5, 4), rng.uniform(0.9, 1.0)
        fast, full = bb._wavefront_contour(dist, 1e-3, ne), bb._wavefront_contour(dist, 1e-3, ne, banded=False)
        assert fast.x.tobytes() == full.x.tobytes() and fast.y.tobytes() == full.y.tobytes()


@pytest.mark.parametrize("tag, gouy, nfront", [("gouy0", 0, 1000), ("nfront500", 1, 500)])
def test_wavefront_class_switches_match_reference(beam_fix, tag, gouy, nfront):
    """Beam._gouy and Beam._nfront (beam.py:23, :27) are class-level settings the finder must honour."""
    from rayrtrace_b200.beam import Beam
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    Beam._gouy, Beam._nfront = gouy, nfront
    try:
        wf = _beam(beam_fix).wavefront_r(dist=3, ampbnd=1e-3, n_field=nf)
    finally:
        Beam._gouy, Beam._nfront = 1, 1000
    assert wf.x.tobytes() == beam_fix[f"{tag}_wf_x"].tobytes()
    assert wf.y.tobytes() == beam_fix[f"{tag}_wf_y"].tobytes()


def test_newton_front_is_the_exact_iso_phase_line(beam_fix):
    """front_method='newton': constant phase to 1e-9 rad, through (0, dist), within the
    marching-squares interpolation error (~1e-6 cm) of the reference's front."""
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    b = _beam(beam_fix)
    b.front_method = "newton"
    wf = b.wavefront_r(dist=3, ampbnd=1e-3, n_field=nf)
    ne = nf.refractive_index(freq=b.freq)
    ph = b._field(wf.x, wf.y, n=ne, outfmt="dict")["phase"]
    assert np.ptp(ph) < 1e-9
    mid = len(wf.x) // 2
    assert abs(wf.y[mid] - 3) < b.wave_len          # nearest whole-period front to dist = 3 (beam.py:263-272)
    assert abs(ph[mid] / 2 / np.pi - round(ph[mid] / 2 / np.pi)) < 1e-6
    ref = interp1d(beam_fix["wf_x"], beam_fix["wf_y"], kind="cubic")
    xs = np.linspace(wf.x[1], wf.x[-2], 4001)
    assert 1e-8 < np.abs(wf(xs) - ref(xs)).max() < 3e-6


def test_launch_conditions_match_reference_beam_rays(beam_fix):
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    b = _beam(beam_fix)
    wf = interp1d(beam_fix["wf_x"], beam_fix["wf_y"], kind="cubic")
    got = b.launch_conditions(int(beam_fix["rnum"]), wf, nf.refractive_index(freq=b.freq))
    want = beam_fix["launch"]
    for g, w, name in zip(got, want, ("x", "y", "kx", "ky", "amp", "phase")):
        np.testing.assert_allclose(g, w, rtol=1e-13, atol=1e-13, err_msg=name)


@pytest.mark.parametrize("rnum", [7, 250, 1000])
def test_launch_conditions_for_other_ray_counts(beam_fix, rnum):
    """The reference takes the launch normals from np.gradient on a fine grid of lcm(rnum, 1000) points
    (beam.py:306-326); the product evaluates only the 3 * rnum nodes that enter.  Coprime, divisor
    and equal ray counts all reproduce the reference's launch set."""
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    b = _beam(beam_fix)
    wf = b.wavefront_r(dist=3, ampbnd=1e-3, n_field=nf)
    got = b.launch_conditions(rnum, wf, nf.refractive_index(freq=b.freq))
    for g, w, name in zip(got, beam_fix[f"launch_r{rnum}"], ("x", "y", "kx", "ky", "amp", "phase")):
        np.testing.assert_allclose(g, w, rtol=1e-12, atol=1e-12, err_msg=name)


def _host_ray(oracle_mod, nf, launch, x0, y0, L=4096):
    """A product Ray filled from an ORACLE trajectory (test only: no GPU here)."""
    from rayrtrace_b200.ray import Ray
    tr = oracle_mod.trace_batch(*[np.array([v]) for v in launch], L, nf.st, nf.dx[0], nf.dy[0], nf.nx, nf.ny, nf.ax, nf.ay)
    k = int(tr["nsteps"][0])
    t = tr["traj"][0, :k]
    r = Ray.__new__(Ray)
    r.n_field, r.amp, r.freq = nf, 1.0, 136.268e9
    r.x = x0 + (t[:, 0] - t[0, 0]) * nf.dx[0]
    r.y = y0 + (t[:, 1] - t[0, 1]) * nf.dy[0]
    r.kx, r.ky = t[:, 2].copy(), t[:, 3].copy()
    r.st = np.full_like(r.x, nf.st)
    r.st[-1] = 0
    return r


def test_ray_addpoint_matches_reference(oracle_mod, ft2_cross):
    """rayrtrace_b200.ray.Ray.addpoint against what the reference's Ray.addpoint returned."""
    n, x, y = ft2_density(int(ft2_cross["nx"]), float(ft2_cross["n0"]))
    nf = Nfield(n, x=x, y=y)
    res = ft2_cross["res"]
    checked = 0
    for li, (val, axis) in enumerate(zip(ft2_cross["line_val"], ft2_cross["line_axis"])):
        for r in range(0, ft2_cross["launch"].shape[0], 2):
            ray = _host_ray(oracle_mod, nf, ft2_cross["launch"][r], ft2_cross["x0"][r], ft2_cross["y0"][r])
            n0 = len(ray.x)
            want = res[li, r]
            if want[8] == 1:                             # the reference raised (no non-negative root)
                with pytest.raises(ValueError):
                    ray.addpoint(float(val), axis=str(axis))
                continue
            ind = ray.addpoint(float(val), axis=str(axis))
            if want[0] == 0:
                assert not hasattr(ind, "__len__") or len(ind) == 0
                continue
            assert len(ind) == int(want[0]) and int(ind[0]) == int(want[1])
            assert len(ray.x) == n0 + len(ind) == len(ray.st) == len(ray.kx)
            i = int(ind[0])
            np.testing.assert_allclose([ray.x[i], ray.y[i], ray.kx[i], ray.ky[i], ray.lax[i], ray.ph[i]], want[2:8],
                                       rtol=1e-11, atol=1e-11)
            checked += 1
    assert checked > 40


def test_pages_vectorised_equals_reference_loop():
    """beam._pages (one vectorised pass per page) == the literal loop of beam.py:374-398 on random
    receiver coordinates: folded beams, ties, duplicates, reversed halves, empty and tiny inputs."""
    from rayrtrace_b200.beam import _pages, _pages_loop
    rng = np.random.default_rng(0)
    for trial in range(1500):
        n = int(rng.integers(0, 70))
        kind = trial % 5
        if kind == 0:
            x = rng.normal(size=n)
        elif kind == 1:
            x = np.sort(rng.normal(size=n)) + rng.normal(size=n) * 0.2
        elif kind == 2:
            x = rng.integers(0, 6, size=n).astype(float)
        elif kind == 3:
            x = np.concatenate([np.sort(rng.normal(size=n // 2)), np.sort(rng.normal(size=n - n // 2))[::-1]])
        else:
            x = np.sort(rng.normal(size=n))
        a, b = _pages(x), _pages_loop(x)
        assert len(a) == len(b) and all(np.array_equal(p, q) for p, q in zip(a, b)), (trial, x)
