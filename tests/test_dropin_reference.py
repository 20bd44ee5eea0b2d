"""The UNMODIFIED reference Python layers (ray.py, den.py, beam.py, ft2.py) running on the B200 library
through the drop-in ``sidtrace`` module (rayrtrace_b200/dropin/sidtrace.py) -- SURVEY 8b: "keeps the
Python surface of the sidtrace module as ray.py/beam.py call it".

The reference sources cannot live in this repo and /root/reference does not exist on the GPU box, so the
test runs only when REF_UNMODIFIED_DIR points at a directory holding them.  tools/dropin_check_cmd.py
(run in the build container) prints a gpurun command that ships the four files to /tmp on the box inside
the command line itself and runs this test there; its log is committed under profiles/.

Checked: every array the reference's own objects end up with equals, bit for bit, what the same objects
held when the reference ran on its own C extension (tests/golden/*.npz, made by make_golden.py) and
what the reference's pure-Python twin (ray.py:85-96) computes on the box.
"""
import os
import sys
import warnings

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.gpu

REF_DIR = os.environ.get("REF_UNMODIFIED_DIR", "")
if not (REF_DIR and os.path.exists(os.path.join(REF_DIR, "ray.py"))):
    pytest.skip("REF_UNMODIFIED_DIR not set: the unmodified reference sources are not on this machine",
                allow_module_level=True)


def ft2_density(nx, n0, edge=1e12, r_dia=7.85, half=10.0):
    x = np.linspace(-half, half, nx)
    X, Y = np.meshgrid(x, x)
    return n0 * np.maximum(0, 1 - (X ** 2 + Y ** 2) / r_dia ** 2) + edge, x, x


@pytest.fixture(scope="module")
def refmods():
    import rayrtrace_b200
    from oracle import refenv
    if os.environ.get("DROPIN_TEST_SELFCHECK"):     # build container, no GPU: the test logic on the compiled reference
        import oracle
        dropin = oracle.ref()
    else:
        sys.path.insert(0, rayrtrace_b200.dropin_path())
        import sidtrace as dropin                   # rayrtrace_b200/dropin/sidtrace.py
        assert "rayrtrace_b200" in os.path.realpath(dropin.__file__)
    refenv.install(root=REF_DIR, sidtrace_module=dropin)
    warnings.simplefilter("ignore")
    import den, ray, beam, ft2                      # noqa: E401  the reference's own modules
    for m in (den, ray, beam, ft2):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF_DIR))
    assert ray.trace is dropin.trace                # ray.py:4 bound the B200 entry point
    return den, ray, beam, ft2


def test_reference_ray_on_the_b200_library(refmods):
    den, ray, _, _ = refmods
    g = np.load(os.path.join(GOLDEN, "ft2_ray.npz"))
    n, x, y = ft2_density(int(g["nx"]), float(g["n0"]))
    nf = den.Nfield(n, x=x, y=y)
    ray.Ray.traceType = "c"
    rc = ray.Ray(-3.0, -9.5, 0, 1, float(g["freq"]), 1, n_field=nf)
    for k in ("x", "y", "kx", "ky", "st", "lax", "ph", "k"):
        assert np.array_equal(getattr(rc, k), g["ray_" + k]), k
    ray.Ray.traceType = "py"                        # the reference's own Python loop, on this machine
    rp = ray.Ray(-3.0, -9.5, 0, 1, float(g["freq"]), 1, n_field=nf)
    ray.Ray.traceType = "c"
    for k in ("x", "y", "kx", "ky"):
        assert np.array_equal(getattr(rc, k), getattr(rp, k)), k
    for tag in ("k2", "khalf"):
        ray.Ray.kStep = float(g[tag + "_kstep"])
        rk = ray.Ray(-3.0, -9.5, 0, 1, float(g["freq"]), 1, n_field=nf)
        ray.Ray.kStep = 1
        assert len(rk.x) == int(g[tag + "_n"])
        assert np.array_equal(rk.x, g[tag + "_x"]) and np.array_equal(rk.y, g[tag + "_y"])
        assert rk.ph[-1] == float(g[tag + "_ph_end"])


def test_reference_ray_fan_equals_its_python_twin(refmods):
    den, ray, _, _ = refmods
    n, x, y = ft2_density(129, 3.0e14)              # over-dense: reflected rays
    nf = den.Nfield(n, x=x, y=y)
    rng = np.random.default_rng(5)
    for _ in range(12):
        x0, ang = rng.uniform(-6, 6), rng.uniform(-0.35, 0.35)
        ray.Ray.traceType = "c"
        rc = ray.Ray(x0, -9.5, np.sin(ang), np.cos(ang), 136.268e9, 1, n_field=nf)
        ray.Ray.traceType = "py"
        rp = ray.Ray(x0, -9.5, np.sin(ang), np.cos(ang), 136.268e9, 1, n_field=nf)
        ray.Ray.traceType = "c"
        assert len(rc.x) == len(rp.x)
        for k in ("x", "y", "kx", "ky", "ph"):
            assert np.array_equal(getattr(rc, k), getattr(rp, k)), k


def test_reference_beam_field_result_on_the_b200_library(refmods):
    den, _, beam, _ = refmods
    g = np.load(os.path.join(GOLDEN, "beam.npz"))
    n, x, y = ft2_density(int(g["nx"]), float(g["n0"]))
    nf = den.Nfield(n, x=x, y=y)
    x0, y0 = float(g["x0"]), float(g["y0"])
    tx = beam.Beam(x0=x0, y0=y0, kx0=0, ky0=1)
    rx = beam.Beam(x0=x0, y0=-y0, kx0=0, ky0=-1)
    out = tx.rays(rnum=int(g["rnum"]), dist=3, ampbnd=1e-3, n_field=nf)
    assert np.array_equal(np.stack(out[:6]), g["launch"])
    assert np.array_equal(np.array([len(r.x) for r in tx._rays]), g["nkept"])
    field, xo, x_max, ph_max, signal, _ = beam.field_result(tx, rx)
    assert np.array_equal(field, g["field"]) and np.array_equal(xo, g["xout"])
    assert x_max == g["x_max"] and ph_max == float(g["ph_max"]) and signal == complex(g["signal"])


def test_reference_ft2_phases_on_the_b200_library(refmods):
    _, _, _, ft2 = refmods
    g = np.load(os.path.join(GOLDEN, "ft2_phases.npz"))
    n, x, y = ft2_density(int(g["nx"]), float(g["n0"]))
    ph_max, ph_sig = ft2.phases(n, x, y, a_set=g["a_set"], nray=int(g["nray"]))
    assert np.array_equal(ph_max, g["ph_by_max"]) and np.array_equal(ph_sig, g["ph_by_sig"])
