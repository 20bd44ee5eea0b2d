"""Host-side field producer (rayrtrace_b200.nfield) against the reference's
den.Nfield: committed fixtures always, the live reference when it is present."""
import os

import numpy as np
import pytest

from _util import ft2_density, sha
from rayrtrace_b200.nfield import Nfield, critical_den

HAVE_REF = os.path.exists("/root/reference/den.py") and __import__("oracle").have_ref()


def test_ft2_field_matches_reference_fixture(ft2_ray):
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    assert nf.freq == float(ft2_ray["field_freq"])
    assert float(nf.st).hex() == float(ft2_ray["st"]).hex()
    assert float(nf.dx[0]).hex() == float(ft2_ray["dx0"]).hex()
    assert sha(nf.ax) == str(ft2_ray["ax_sha"])
    assert sha(nf.ay) == str(ft2_ray["ay_sha"])
    assert nf.ax[::20, ::20].tobytes() == ft2_ray["ax_sample"].tobytes()
    assert nf.ay[::20, ::20].tobytes() == ft2_ray["ay_sample"].tobytes()
    assert nf.ax.shape == (nf.ny - 1, nf.nx - 1)


@pytest.mark.parametrize("tag", ["under", "over"])
def test_fan_fields_match_reference_fixture(ft2_fan, tag):
    n, x, y = ft2_density(int(ft2_fan[f"{tag}_nx"]), float(ft2_fan[f"{tag}_n0"]))
    nf = Nfield(n, x=x, y=y)
    assert float(nf.st).hex() == float(ft2_fan[f"{tag}_st"]).hex()
    assert sha(nf.ax) == str(ft2_fan[f"{tag}_ax_sha"]) and sha(nf.ay) == str(ft2_fan[f"{tag}_ay_sha"])


def test_launch_state_matches_reference_fixture(ft2_ray):
    """ray.py:28-68 restated (rayrtrace_b200.ray) gives the launch state the reference Ray used."""
    from rayrtrace_b200.ray import launch_states
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    xg, yg, kx, ky, _ = launch_states(nf, [float(ft2_ray["x0"])], [float(ft2_ray["y0"])], 0.0, 1.0, float(ft2_ray["freq"]))
    want = ft2_ray["launch"]
    assert xg[0] == want[0] and yg[0] == want[1]
    # the cubic interpolant of n is SciPy's spline on both sides; allow 1e-12 relative
    np.testing.assert_allclose([kx[0], ky[0]], want[2:], rtol=1e-12, atol=0)


def test_vacuum_field_and_step():
    nf = Nfield(np.zeros((5, 9)), x=np.linspace(0, 4, 9), y=np.linspace(0, 1, 5))
    assert not nf.ax.any() and not nf.ay.any()
    c = 299792458.0 * 100
    assert nf.st == nf.min_dxy * c / (2 * np.pi * nf.freq)     # den.py:107 with st_grad = inf
    assert nf.refractive_index() == 1.0


def test_bad_shapes_raise():
    with pytest.raises(ValueError):
        Nfield(np.zeros((5, 9)), x=np.arange(4), y=np.arange(5))
    with pytest.raises(ValueError):
        Nfield(np.zeros(5))


def test_critical_density_value():
    # n_c = 2.303e14 cm^-3 at 136.26 GHz
    assert abs(critical_den(136.26e9) / 2.303e14 - 1) < 1e-3


@pytest.mark.skipif(not HAVE_REF, reason="reference sources not present (GPU box)")
def test_against_live_reference_random_density():
    from oracle import refenv
    den, _ = refenv.modules()
    rng = np.random.default_rng(3)
    for ny, nx in ((12, 17), (40, 33)):
        n = rng.uniform(0, 5e13, (ny, nx))
        x = np.cumsum(rng.uniform(0.05, 0.2, nx))     # non-equidistant: gradients use per-cell dx (den.py:62-66)
        y = np.linspace(-3, 3, ny)
        for freq in (136.26e9, 100e9):
            a = den.Nfield(n, x=x, y=y, freq=freq)
            b = Nfield(n, x=x, y=y, freq=freq)
            assert a.ax.tobytes() == b.ax.tobytes() and a.ay.tobytes() == b.ay.tobytes()
            assert float(a.st).hex() == float(b.st).hex()
            assert float(a.x2grid(x[3] + 0.01)) == float(b.x2grid(x[3] + 0.01))
            np.testing.assert_allclose(np.asarray(a.refractive_index(freq, x[4], y[5])).reshape(-1)[0],
                                       np.asarray(b.refractive_index(freq, x[4], y[5])).reshape(-1)[0], rtol=1e-12)
