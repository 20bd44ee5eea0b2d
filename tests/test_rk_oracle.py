"""CPU checks of the adaptive-RK extension's checker (oracle/rk_oracle.c).

PARITY UNPINNED: the reference has no adaptive integrator / interpolated medium / X mode
(SURVEY 8 f-4), so there is no reference output.  The checker is pinned instead by
  * closed-form solutions (vacuum, linear density ramp -> parabola, uniform absorption),
  * an independent SciPy DOP853 integration over independently written interpolants (SciPy's
    natural CubicSpline tensor product for the B-spline mode, a basis-matrix Catmull-Rom patch),
  * the conserved Hamiltonian, the analytic X-mode cutoff, finite-difference derivatives,
  * first-order convergence of the REFERENCE scheme (sidtracemodule.c:36-48 on den.py:65-66
    differences, via the pinned oracle) towards this solution as the grid is refined.
The CUDA kernel is then compared with this checker bit for bit in tests/test_rk_gpu.py.
"""
import numpy as np
import pytest
import scipy.constants as const
from scipy.integrate import solve_ivp
from scipy.interpolate import CubicSpline

import oracle
from oracle import rk
from rayrtrace_b200.nfield import Nfield, critical_den

FREQ = 136.268e9


def grid(n=201, half=10.0):
    xs = np.linspace(-half, half, n)
    return xs, xs[1] - xs[0]


def gaussian_medium(n=201, peak=0.5, width=4.0, **kw):
    xs, d = grid(n)
    XX, YY = np.meshgrid(xs, xs)
    X = peak * np.exp(-(XX ** 2 + YY ** 2) / width ** 2)
    return rk.Medium(X, xs[0], xs[0], d, d, **kw), X, xs, d


# ---- closed forms -----------------------------------------------------------------------------
def test_vacuum_is_a_straight_line():
    xs, d = grid(101)
    m = rk.Medium(np.zeros((101, 101)), xs[0], xs[0], d, d)
    th = 0.7
    r = rk.trace(m, [-9.0], [-9.0], [np.cos(th)], [np.sin(th)], tau_max=11.0, hmax=1.0, L=64, aux=True)
    assert r["status"][0] == rk.TAU_MAX and r["nrej"][0] == 0
    assert r["tau"][0] == 11.0
    np.testing.assert_allclose(r["end_state"][:, 0], [-9 + 11 * np.cos(th), -9 + 11 * np.sin(th), np.cos(th), np.sin(th)],
                               rtol=0, atol=2e-14)
    assert r["phi"][0] == pytest.approx(11.0, abs=1e-13)       # |p| = 1: optical path = tau
    assert abs(r["ham"][0]) < 1e-15 and r["n2_min"][0] == 1.0
    k = r["nacc"][0]
    assert np.all(np.isfinite(r["traj"][0, :k + 1])) and np.all(np.isnan(r["traj"][0, k + 1:]))
    np.testing.assert_array_equal(r["aux"][0, k], [11.0, 0.0])


@pytest.mark.parametrize("interp", [rk.BILINEAR, rk.CATMULL_ROM, rk.BSPLINE])
def test_linear_ramp_is_a_parabola(interp):
    """X = a + b x, O mode: dp_x/dtau = -b/2 exactly; every interpolant reproduces a linear table."""
    xs, d = grid(201)
    a, b = 0.4, 0.04
    X = np.tile(a + b * xs, (201, 1))
    m = rk.Medium(X, xs[0], xs[0], d, d, interp=interp)
    x0, y0 = -9.5, -9.5
    N = np.sqrt(1 - (a + b * x0))
    px, py = N * np.cos(0.9), N * np.sin(0.9)
    T = 20.0
    r = rk.trace(m, [x0], [y0], [px], [py], tau_max=T, hmax=0.5)
    assert r["status"][0] == rk.TAU_MAX
    want = [x0 + px * T - 0.25 * b * T ** 2, y0 + py * T, px - 0.5 * b * T, py]
    np.testing.assert_allclose(r["end_state"][:, 0], want, rtol=0, atol=5e-12)
    assert abs(r["ham"][0]) < 1e-12


def test_uniform_absorption_and_deposited_power():
    m, X, xs, d = gaussian_medium(alpha=np.full((201, 201), 0.03))
    n = 64
    x0 = np.linspace(-4, 4, n)
    amp = np.linspace(0.5, 1.5, n)
    r = rk.trace(m, x0, np.full(n, -9.5), np.zeros(n), np.ones(n), amp=amp, nbins=40, r_lcs=7.85, rho_max=2.0,
                 hmax=0.4, threads=4)
    np.testing.assert_allclose(r["sig"], 0.03 * r["tau"], rtol=1e-12)
    absorbed = amp ** 2 * (1 - np.exp(-r["sig"]))
    # every step's midpoint lies within rho < 2 here, so all absorbed power is binned
    assert r["hist"].sum() == pytest.approx(absorbed.sum(), rel=1e-12)
    assert np.all(r["hist"] >= 0) and r["hist"][:30].min() > 0


# ---- independent interpolation and integration -------------------------------------------------
_CR = 0.5 * np.array([[0, 2, 0, 0], [-1, 0, 1, 0], [2, -5, 4, -1], [-1, 3, -3, 1.0]])   # rows: 1, t, t^2, t^3


def _catmull_rom(X, org, d, x, y):
    """Catmull-Rom patch written from its basis matrix (not from the kernel's Horner weights)."""
    ny, nx = X.shape
    u, v = (x - org) / d, (y - org) / d
    iu, iv = int(np.clip(np.floor(u), 0, nx - 2)), int(np.clip(np.floor(v), 0, ny - 2))
    t, s = u - iu, v - iv
    P = X[np.ix_(np.clip(np.arange(iv - 1, iv + 3), 0, ny - 1), np.clip(np.arange(iu - 1, iu + 3), 0, nx - 1))]
    T, dT = np.array([1, t, t * t, t ** 3]), np.array([0, 1, 2 * t, 3 * t * t])
    S, dS = np.array([1, s, s * s, s ** 3]), np.array([0, 1, 2 * s, 3 * s * s])
    wx, dwx, wy, dwy = T @ _CR, dT @ _CR, S @ _CR, dS @ _CR
    return wy @ P @ wx, (wy @ P @ dwx) / d, (dwy @ P @ wx) / d


class _NaturalSpline:
    """Tensor product of SciPy's natural cubic splines: what the prefiltered B-spline must equal."""

    def __init__(self, X, xs, ys):
        self.rows = CubicSpline(xs, X, axis=1, bc_type="natural")
        self.ys = ys

    def __call__(self, x, y):
        col = CubicSpline(self.ys, self.rows(x), bc_type="natural")
        dcol = CubicSpline(self.ys, self.rows(x, 1), bc_type="natural")
        return float(col(y)), float(dcol(y)), float(col(y, 1))


def test_prefiltered_bspline_is_the_natural_cubic_spline():
    xs = np.linspace(-10, 10, 101)
    ys = np.linspace(-8, 8, 81)
    XX, YY = np.meshgrid(xs, ys)
    F = 0.5 * np.exp(-(XX ** 2 + YY ** 2) / 16) * (1 + 0.1 * np.sin(XX))
    m = rk.Medium(F, xs[0], ys[0], xs[1] - xs[0], ys[1] - ys[0])
    ref = _NaturalSpline(F, xs, ys)
    rng = np.random.default_rng(0)
    for _ in range(100):
        x, y = rng.uniform(-10, 10), rng.uniform(-8, 8)
        f, fx, fy = ref(x, y)
        gx, gy, _, n2 = m.eval(x, y)
        assert 1 - n2 == pytest.approx(f, abs=1e-13)
        assert -2 * gx == pytest.approx(fx, abs=1e-12) and -2 * gy == pytest.approx(fy, abs=1e-12)
    # interpolation property at the nodes, and tiny tables
    for j in range(0, 81, 9):
        for i in range(0, 101, 7):
            assert 1 - m.eval(xs[i], ys[j])[3] == pytest.approx(F[j, i], abs=1e-14)
    for shape in ((2, 2), (2, 3), (3, 2), (3, 5)):
        G = rng.uniform(0, 0.5, shape)
        c = rk.prefilter(G)
        assert c.shape == (shape[0] + 2, shape[1] + 2)
        back = (c[:-2] + 4 * c[1:-1] + c[2:]) / 6
        back = (back[:, :-2] + 4 * back[:, 1:-1] + back[:, 2:]) / 6
        np.testing.assert_allclose(back, G, rtol=0, atol=1e-15)


def test_against_scipy_dop853_o_mode_bspline():
    m, X, xs, d = gaussian_medium()
    ref = _NaturalSpline(X, xs, xs)
    x0, y0 = -3.0, -9.5
    N = np.sqrt(1 - ref(x0, y0)[0])

    def rhs(t, z):
        _, Fx, Fy = ref(z[0], z[1])
        return [z[2], z[3], -0.5 * Fx, -0.5 * Fy]

    T = 14.0
    sol = solve_ivp(rhs, [0, T], [x0, y0, 0.0, N], method="DOP853", rtol=1e-12, atol=1e-13, max_step=0.1)
    r = rk.trace(m, [x0], [y0], [0.0], [N], rtol=1e-12, atol_r=1e-12, atol_p=1e-12, tau_max=T, hmax=0.4)
    np.testing.assert_allclose(r["end_state"][:, 0], sol.y[:, -1], rtol=0, atol=1e-8)
    assert abs(r["ham"][0]) < 1e-8


def test_against_scipy_dop853_x_mode_catmull_rom():
    Y0, R0 = 0.45, 55.0
    m, X, xs, d = gaussian_medium(peak=0.35, mode="X", Y0=Y0, R0=R0, interp=rk.CATMULL_ROM)
    x0, y0 = -9.5, 1.0

    def n2_grad(x, y):
        F, Fx, Fy = _catmull_rom(X, xs[0], d, x, y)
        Y2 = (Y0 * R0 / (R0 + x)) ** 2
        D = 1 - F - Y2
        n2 = 1 - F * (1 - F) / D
        dX = -((1 - 2 * F) * D + F * (1 - F)) / D ** 2
        dY2 = -F * (1 - F) / D ** 2
        return n2, dX * Fx + dY2 * (-2 * Y2 / (R0 + x)), dX * Fy

    N = np.sqrt(n2_grad(x0, y0)[0])

    def rhs(t, z):
        _, gx, gy = n2_grad(z[0], z[1])
        return [z[2], z[3], 0.5 * gx, 0.5 * gy]

    T = 12.0
    sol = solve_ivp(rhs, [0, T], [x0, y0, N, 0.0], method="DOP853", rtol=1e-12, atol=1e-13, max_step=0.05)
    r = rk.trace(m, [x0], [y0], [N], [0.0], rtol=1e-12, atol_r=1e-12, atol_p=1e-12, tau_max=T, hmax=0.05)
    np.testing.assert_allclose(r["end_state"][:, 0], sol.y[:, -1], rtol=0, atol=5e-8)
    # and the checker's N^2 at the launch point equals the independent formula
    assert m.eval(x0, y0)[3] == pytest.approx(N * N, rel=1e-14)


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("interp", [rk.BILINEAR, rk.CATMULL_ROM, rk.BSPLINE])
def test_force_is_half_the_gradient_of_n2(mode, interp):
    """Analytic derivatives of the dispersion relation vs central differences of N^2 (inside one
    cell, where every interpolant is smooth)."""
    m, X, xs, d = gaussian_medium(peak=0.4, mode=mode, Y0=0.45, R0=55.0, interp=interp)
    rng = np.random.default_rng(3)
    for _ in range(50):
        i, j = rng.integers(20, 180, 2)
        x, y = xs[i] + d * rng.uniform(0.2, 0.8), xs[j] + d * rng.uniform(0.2, 0.8)
        xsc, ysc = rng.uniform(0.5, 1.5), rng.uniform(0.8, 1.2)
        fx, fy, _, _ = m.eval(x, y, xsc, ysc)
        e = 1e-5
        gx = (m.eval(x + e, y, xsc, ysc)[3] - m.eval(x - e, y, xsc, ysc)[3]) / (2 * e)
        gy = (m.eval(x, y + e, xsc, ysc)[3] - m.eval(x, y - e, xsc, ysc)[3]) / (2 * e)
        assert fx == pytest.approx(0.5 * gx, rel=2e-6, abs=1e-9)
        assert fy == pytest.approx(0.5 * gy, rel=2e-6, abs=1e-9)


# ---- physics -----------------------------------------------------------------------------------
def test_hamiltonian_is_conserved_and_tolerance_controls_the_error():
    m, X, xs, d = gaussian_medium(peak=0.9)
    n = 33
    x0 = np.linspace(-5, 5, n)
    N = np.sqrt([m.eval(x, -9.5)[3] for x in x0])
    worst, steps = [], []
    for tol in (1e-6, 1e-9, 1e-12):
        r = rk.trace(m, x0, np.full(n, -9.5), np.zeros(n), N, rtol=tol, atol_r=tol, atol_p=tol, hmax=1.0, threads=4)
        assert np.all(r["status"] == rk.LEFT_DOMAIN)
        worst.append(np.abs(r["ham"]).max())
        steps.append(int(r["nacc"].sum()))
        assert worst[-1] < 3000 * tol
    assert worst[1] < worst[0] / 30 and worst[2] < worst[1] / 30
    assert steps[0] < steps[1] < steps[2]


def test_o_mode_cutoff_reflects_normal_incidence():
    """Slab X = (x + 10)/10: a ray launched along +x turns where N^2 = 0, i.e. at x = 0."""
    xs, d = grid(201)
    X = np.tile((xs + 10) / 10.0, (201, 1))
    m = rk.Medium(X, xs[0], xs[0], d, d)
    x0 = -9.0
    N = np.sqrt(1 - (x0 + 10) / 10)
    r = rk.trace(m, [x0], [0.0], [N], [0.0], hmax=0.2, L=4000)
    k = r["nacc"][0]
    assert r["status"][0] == rk.LEFT_DOMAIN and r["end_state"][0, 0] < -10 and r["end_state"][2, 0] < 0
    # closed form: x(tau) = x0 + N tau - tau^2/40, deepest point x0 + 10 N^2 = 0
    assert r["traj"][0, :k + 1, 0].max() == pytest.approx(0.0, abs=2e-3)
    assert r["n2_min"][0] == pytest.approx(0.0, abs=2e-4)


def test_x_mode_right_hand_cutoff():
    """Uniform Y (R0 -> inf): normal incidence turns at X = 1 - Y, before the O-mode cutoff."""
    xs, d = grid(401)
    X = np.tile((xs + 10) / 10.0, (401, 1))
    Y = 0.4
    m = rk.Medium(X, xs[0], xs[0], d, d, mode="X", Y0=Y, R0=1e12)
    x0 = -9.5
    n2 = m.eval(x0, 0.0)[3]
    r = rk.trace(m, [x0], [0.0], [np.sqrt(n2)], [0.0], hmax=0.1, L=4000)
    k = r["nacc"][0]
    assert r["status"][0] == rk.LEFT_DOMAIN and r["end_state"][2, 0] < 0
    x_cut = 10 * (1 - Y) - 10
    assert r["traj"][0, :k + 1, 0].max() == pytest.approx(x_cut, abs=2e-3)
    assert abs(r["ham"][0]) < 1e-7


def test_upper_hybrid_resonance_retires_with_step_underflow():
    """From the dense side N^2 -> +inf at X = 1 - Y^2: the step collapses, the ray is flagged."""
    xs, d = grid(401)
    X = np.tile((xs + 10) / 10.0, (401, 1))          # X = 1 at x = 0, resonance (Y = 0.4) at x = -1.6
    m = rk.Medium(X, xs[0], xs[0], d, d, mode="X", Y0=0.4, R0=1e12)
    x0 = 0.0
    n2 = m.eval(x0, 0.0)[3]
    assert n2 > 0
    r = rk.trace(m, [x0], [0.0], [-np.sqrt(n2)], [0.0], hmax=0.1, hmin=1e-10, max_steps=20000)
    assert r["status"][0] == rk.STEP_UNDERFLOW
    assert r["end_state"][0, 0] == pytest.approx(-1.6, abs=1e-3)
    assert r["nacc"][0] > 50 and r["nrej"][0] >= 1
    assert abs(r["end_state"][2, 0]) > 30          # |p| = N blows up


def test_turning_point_of_a_reflected_ray():
    """X = 0.2 + 0.08 (y + 10): p_y falls linearly, the ray turns at y0 + p_y0^2 / b; turn_index is the
    first accepted state with p_y <= 0 (the parity kernel's definition), turn_state that state."""
    xs, d = grid(201)
    b = 0.08
    X = np.tile((0.2 + b * (xs + 10))[:, None], (1, 201))
    m = rk.Medium(X, xs[0], xs[0], d, d)
    y0 = -9.0
    N = np.sqrt(1 - (0.2 + b * (y0 + 10)))
    th = np.array([0.0, 0.3, 0.6])
    r = rk.trace(m, np.full(3, -2.0), np.full(3, y0), N * np.sin(th), N * np.cos(th), hmax=0.2, L=2000)
    y_turn = y0 + (N * np.cos(th)) ** 2 / b
    assert np.all(r["turn_index"] > 0) and np.all(r["status"] == rk.LEFT_DOMAIN)
    for i in range(3):
        k = r["turn_index"][i]
        np.testing.assert_array_equal(r["turn_state"][:, i], r["traj"][i, k])
        assert r["traj"][i, k - 1, 3] > 0 >= r["traj"][i, k, 3]
        assert abs(r["turn_state"][1, i] - y_turn[i]) < 0.02
        assert r["traj"][i, :r["nacc"][i] + 1, 1].max() == pytest.approx(y_turn[i], abs=0.01)
    # a ray through vacuum never turns
    vac = rk.Medium(np.zeros((201, 201)), xs[0], xs[0], d, d)
    r = rk.trace(vac, [0.0], [-9.0], [0.1], [0.9])
    assert r["turn_index"][0] == -1 and np.all(np.isnan(r["turn_state"][:, 0]))


def test_status_codes():
    m, X, xs, d = gaussian_medium()
    r = rk.trace(m, [-3.0, -3.0, 11.0, -3.0], [-9.5, -9.5, 0.0, -9.5], [0, 0, 1.0, 0], [1.0, 1.0, 0, 1.0],
                 tau_max=5.0, max_steps=100000)
    assert list(r["status"]) == [rk.TAU_MAX, rk.TAU_MAX, rk.START_OUTSIDE, rk.TAU_MAX]
    assert r["nacc"][2] == 0 and r["end_state"][0, 2] == 11.0
    r = rk.trace(m, [-3.0], [-9.5], [0.0], [1.0], max_steps=7)
    assert r["status"][0] == rk.MAX_STEPS and r["nacc"][0] == 7
    r = rk.trace(m, [-3.0], [-9.5], [0.0], [1.0], max_steps=0)
    assert r["status"][0] == rk.MAX_STEPS and r["nacc"][0] == 0
    assert r["n2_min"][0] == m.eval(-3.0, -9.5)[3]
    r = rk.trace(m, [np.nan], [-9.5], [0.0], [1.0])
    assert r["status"][0] == rk.START_OUTSIDE


def test_per_ray_scales_equal_scaled_tables():
    m, X, xs, d = gaussian_medium(peak=0.3, mode="X", Y0=0.45, R0=55.0)
    m2 = rk.Medium(1.7 * X, xs[0], xs[0], d, d, mode="X", Y0=0.45 * 0.9, R0=55.0)
    a = rk.trace(m, [-2.0], [-9.5], [0.0], [0.9], xscale=[1.7], yscale=[0.9], hmax=0.3)
    b = rk.trace(m2, [-2.0], [-9.5], [0.0], [0.9], hmax=0.3)
    np.testing.assert_allclose(a["end_state"], b["end_state"], rtol=0, atol=1e-9)


def test_threads_do_not_change_results():
    m, X, xs, d = gaussian_medium(alpha=np.full((201, 201), 0.01))
    n = 257
    x0 = np.linspace(-6, 6, n)
    a = rk.trace(m, x0, np.full(n, -9.5), np.zeros(n), np.ones(n), nbins=16, r_lcs=7.85, rho_max=1.5, threads=1)
    b = rk.trace(m, x0, np.full(n, -9.5), np.zeros(n), np.ones(n), nbins=16, r_lcs=7.85, rho_max=1.5, threads=5)
    for k in ("end_state", "tau", "phi", "sig", "nacc", "nrej", "status"):
        np.testing.assert_array_equal(a[k], b[k])
    np.testing.assert_allclose(a["hist"], b["hist"], rtol=1e-12)


# ---- stop line (receiver) ----------------------------------------------------------------------
def test_stop_line_lands_on_the_line_closed_forms():
    xs, d = grid(201)
    vac = rk.Medium(np.zeros((201, 201)), xs[0], xs[0], d, d)
    th = np.linspace(-0.4, 0.4, 11)
    r = rk.trace(vac, np.full(11, 0.5), np.full(11, -9.5), np.sin(th), np.cos(th), hmax=0.4, stop=(8.0, "y"))
    assert np.all(r["status"] == rk.STOP_LINE)
    assert np.abs(r["end_state"][1] - 8.0).max() <= 1e-12
    np.testing.assert_allclose(r["tau"], 17.5 / np.cos(th), rtol=1e-14)
    np.testing.assert_allclose(r["end_state"][0], 0.5 + 17.5 * np.tan(th), rtol=0, atol=1e-13)
    np.testing.assert_allclose(r["phi"], r["tau"], rtol=1e-14)
    # linear ramp: the parabola's crossing of y = 2 and of x = -5
    X = np.tile(0.4 + 0.04 * xs, (201, 1))
    ramp = rk.Medium(X, xs[0], xs[0], d, d)
    N = np.sqrt(1 - (0.4 + 0.04 * -9.5))
    px, py = N * np.cos(0.5), N * np.sin(0.5)
    r = rk.trace(ramp, [-9.5], [-3.0], [px], [py], hmax=0.5, stop=(2.0, "y"))
    T = 5.0 / py
    assert r["status"][0] == rk.STOP_LINE and abs(r["tau"][0] - T) < 1e-13
    assert abs(r["end_state"][0, 0] - (-9.5 + px * T - 0.01 * T * T)) < 1e-13
    r = rk.trace(ramp, [-9.5], [-3.0], [px], [py], hmax=0.5, stop=(-5.0, "x"))
    T = (px - np.sqrt(px * px - 4 * 0.01 * 4.5)) / 0.02          # -9.5 + px T - 0.01 T^2 = -5, first root
    assert r["status"][0] == rk.STOP_LINE and abs(r["end_state"][0, 0] + 5.0) <= 1e-12
    assert abs(r["tau"][0] - T) < 1e-12
    # a ray that never reaches the line ends as usual; one that starts on it is not stopped there
    r = rk.trace(vac, [-3.0, -3.0], [-9.5, 8.0], [1.0, 0.0], [0.0, 1.0], hmax=0.4, stop=(8.0, "y"))
    assert list(r["status"]) == [rk.LEFT_DOMAIN, rk.LEFT_DOMAIN]


def test_stop_line_in_a_plasma_agrees_with_scipy_events():
    m, X, xs, d = gaussian_medium(peak=0.4)
    ref = _NaturalSpline(X, xs, xs)
    x0, y0 = -1.0, -9.5
    N = np.sqrt(1 - ref(x0, y0)[0])

    def rhs(t, z):
        _, Fx, Fy = ref(z[0], z[1])
        return [z[2], z[3], -0.5 * Fx, -0.5 * Fy]

    def hit(t, z):
        return z[1] - 8.0
    hit.terminal, hit.direction = True, 1
    sol = solve_ivp(rhs, [0, 60], [x0, y0, 0.1 * N, np.sqrt(1 - 0.01) * N], method="DOP853", rtol=1e-12, atol=1e-13,
                    max_step=0.1, events=hit)
    r = rk.trace(m, [x0], [y0], [0.1 * N], [np.sqrt(1 - 0.01) * N], rtol=1e-12, atol_r=1e-12, atol_p=1e-12, hmax=0.4,
                 stop=(8.0, "y"))
    assert r["status"][0] == rk.STOP_LINE and abs(r["end_state"][1, 0] - 8.0) <= 1e-12
    assert r["tau"][0] == pytest.approx(sol.t_events[0][0], abs=2e-8)
    np.testing.assert_allclose(r["end_state"][:, 0], sol.y_events[0][0], rtol=0, atol=2e-8)


def test_reference_receiver_phase_converges_to_the_rk_phase():
    """The reference's eikonal phase at a receiver line (Ray.ph after Ray.addpoint, ray.py:113-224,
    on the trajectory of the pinned oracle) approaches k0 * (optical path of the RK ray landed on the
    same line) at first order in the grid spacing."""
    from rayrtrace_b200.ray import Ray
    nc = critical_den(FREQ)
    k0 = 2 * np.pi * FREQ / (const.speed_of_light * 100)
    x0, y0, yrx = -3.0, -9.5, 8.0
    N0 = np.sqrt(1 - 0.5 * np.exp(-(x0 ** 2 + y0 ** 2) / 16))
    diffs, rk_phase = [], []
    for n in (201, 401, 801):
        xs, d = grid(n)
        XX, YY = np.meshgrid(xs, xs)
        den = 0.5 * nc * np.exp(-(XX ** 2 + YY ** 2) / 16.0)
        nf = Nfield(den, x=xs, y=xs, freq=FREQ)
        L = 40 * n
        tr = oracle.trace_batch(np.array([(x0 + 10) / d]), np.array([(y0 + 10) / d]), np.array([0.0]), np.array([k0 * N0]),
                                L, nf.st, nf.dx[0], nf.dy[0], n, n, nf.ax, nf.ay)
        t = tr["traj"][0, :int(tr["nsteps"][0])]
        ray = Ray.__new__(Ray)
        ray.n_field, ray.amp, ray.freq = nf, 1.0, FREQ
        ray.x = x0 + (t[:, 0] - t[0, 0]) * nf.dx[0]
        ray.y = y0 + (t[:, 1] - t[0, 1]) * nf.dy[0]
        ray.kx, ray.ky = t[:, 2].copy(), t[:, 3].copy()
        ray.st = np.full_like(ray.x, nf.st)
        ray.st[-1] = 0
        idx = ray.addpoint(yrx, "y")
        ph_ref = ray.ph[int(np.atleast_1d(idx)[0])]
        m = rk.Medium(den / nc, xs[0], xs[0], d, d)
        q = rk.trace(m, [x0], [y0], [0.0], [N0], rtol=1e-11, atol_r=1e-11, atol_p=1e-11, hmax=0.2, stop=(yrx, "y"))
        assert q["status"][0] == rk.STOP_LINE
        rk_phase.append(k0 * q["phi"][0])
        diffs.append(abs(ph_ref - rk_phase[-1]))
    assert max(rk_phase) - min(rk_phase) < 2e-3                 # rad, of ~400: the RK phase barely depends on the grid
    assert diffs[0] < 1.0 and 1.6 < diffs[0] / diffs[1] < 2.6 and 1.6 < diffs[1] / diffs[2] < 2.6


# ---- the reference scheme converges to this solution ---------------------------------------
def _line_crossing(xy, yline):
    x, y = xy[:, 0], xy[:, 1]
    j = np.nonzero((y[:-1] < yline) & (y[1:] >= yline))[0][0]
    f = (yline - y[j]) / (y[j + 1] - y[j])
    return x[j] + f * (x[j + 1] - x[j])


def test_reference_scheme_converges_to_the_rk_solution_at_first_order():
    """sidtracemodule.c:36-48 on the forward differences of den.py:65-66 is a first-order scheme:
    its crossing of the line y = 8 approaches the adaptive-RK one like dx."""
    nc = critical_den(FREQ)
    k0 = 2 * np.pi * FREQ / (const.speed_of_light * 100)
    x0, y0 = -3.0, -9.5
    N0 = np.sqrt(1 - 0.5 * np.exp(-(x0 ** 2 + y0 ** 2) / 16))
    diffs, rk_cross = [], []
    for n in (201, 401, 801):
        xs, d = grid(n)
        XX, YY = np.meshgrid(xs, xs)
        den = 0.5 * nc * np.exp(-(XX ** 2 + YY ** 2) / 16.0)
        nf = Nfield(den, x=xs, y=xs, freq=FREQ)
        L = 40 * n
        r = oracle.trace_batch(np.array([(x0 + 10) / d]), np.array([(y0 + 10) / d]), np.array([0.0]), np.array([k0 * N0]),
                               L, nf.st, nf.dx[0], nf.dy[0], n, n, nf.ax, nf.ay)
        tr = r["traj"][0, :r["nsteps"][0] + 1]
        xy = np.stack([x0 + (tr[:, 0] - tr[0, 0]) * nf.dx[0], y0 + (tr[:, 1] - tr[0, 1]) * nf.dy[0]], 1)
        m = rk.Medium(den / nc, xs[0], xs[0], d, d)
        q = rk.trace(m, [x0], [y0], [0.0], [N0], rtol=1e-10, atol_r=1e-10, atol_p=1e-10, hmax=0.05, h0=0.01, L=4000)
        T = q["traj"][0, :q["nacc"][0] + 1]
        rk_cross.append(_line_crossing(T[:, :2], 8.0))
        diffs.append(abs(_line_crossing(xy, 8.0) - rk_cross[-1]))
    assert max(rk_cross) - min(rk_cross) < 5e-5            # the RK answer barely depends on the grid
    assert diffs[0] < 0.08 and diffs[2] < 0.02
    assert 1.7 < diffs[0] / diffs[1] < 2.4 and 1.7 < diffs[1] / diffs[2] < 2.4


# ---------------------------------------------------------------- X mode with |B| from a tabulated equilibrium (mode "XT")
def _solovev_tables(nr=257, nz=257, half=0.12, R0=0.55, a=0.08, kappa=1.2, B0=2.2, psi_b=6e-3, nf=33):
    R = np.linspace(R0 - half, R0 + half, nr)
    Z = np.linspace(-half, half, nz)
    RR, ZZ = np.meshgrid(R, Z, indexing="ij")
    c = psi_b / (a * a * R0 * R0)
    psi = c * (RR ** 2 * ZZ ** 2 / kappa ** 2 + (RR ** 2 - R0 ** 2) ** 2 / 4)
    eq = dict(psi=psi, fpol=np.full(nf, B0 * R0), rleft=R0 - half, rdim=2 * half, zmid=0.0, zdim=2 * half, simagx=0.0,
              sibdry=psi_b, r0_cm=100 * R0, z0_cm=0.0, y_per_tesla=2.799249e10 / 60e9)

    def B(Rm, Zm):
        return -c * 2 * Rm * Zm / kappa ** 2, c * (2 * Zm ** 2 / kappa ** 2 + (Rm ** 2 - R0 ** 2)), B0 * R0 / Rm
    return eq, B


def test_equilibrium_table_is_the_solovev_field():
    """|B|^2 from the psi(R, Z) table (the reference's Equilibrium fields, equilibrium.py:11-49): the toroidal part is
    exact, the poloidal part (central differences at the nodes, interpolated bilinearly: a continuous field)
    converges at second order in the psi grid spacing."""
    xs = np.linspace(-10, 10, 101)
    errs = []
    for n in (65, 129, 257):
        eq, B = _solovev_tables(nr=n, nz=n)
        y2 = rk.equilibrium_y2(eq, xs.size, xs.size, xs[0], xs[0], xs[1] - xs[0], xs[1] - xs[0])
        RR, ZZ = (55.0 + xs[None, :]) * 0.01 + 0 * xs[:, None], xs[:, None] * 0.01 + 0 * xs[None, :]
        BR, BZ, Bp = B(RR, ZZ)
        exact = eq["y_per_tesla"] ** 2 * (BR ** 2 + BZ ** 2 + Bp ** 2)
        pol = eq["y_per_tesla"] ** 2 * (BR ** 2 + BZ ** 2)
        errs.append(np.max(np.abs(y2 - exact)) / pol.max())
    assert errs[2] < 5e-5 and errs[0] / errs[2] > 12.0       # second order: 4x finer grid, ~16x smaller error
    assert np.max(np.abs(y2 - exact) / exact) < 5e-6


def test_tabulated_field_force_is_the_gradient_of_n2():
    """Mode XT: fx, fy = grad(N^2)/2 with N^2 = 1 - X(1-X)/(1-X-Y^2) and BOTH X and Y^2 interpolated: centred
    differences of the N^2 the evaluation itself reports."""
    xs = np.linspace(-10, 10, 161)
    X = 0.3 * np.exp(-((xs[None, :] - 0.5) ** 2 + 1.2 * xs[:, None] ** 2) / 15.0)
    eq, _ = _solovev_tables()
    d = xs[1] - xs[0]
    y2 = rk.equilibrium_y2(eq, xs.size, xs.size, xs[0], xs[0], d, d)
    med = rk.Medium(X, xs[0], xs[0], d, d, mode="XT", Y2=y2)
    rng = np.random.default_rng(5)
    for _ in range(40):
        x, y = rng.uniform(-8, 8, 2)
        xs_, ys_ = rng.uniform(0.5, 1.5), rng.uniform(0.8, 1.2)
        fx, fy, _, n2 = med.eval(x, y, xs_, ys_)
        h = 1e-5
        gx = (med.eval(x + h, y, xs_, ys_)[3] - med.eval(x - h, y, xs_, ys_)[3]) / (2 * h)
        gy = (med.eval(x, y + h, xs_, ys_)[3] - med.eval(x, y - h, xs_, ys_)[3]) / (2 * h)
        assert abs(fx - gx / 2) < 1e-7 * max(1, abs(gx)) and abs(fy - gy / 2) < 1e-7 * max(1, abs(gy))


def test_tabulated_toroidal_field_equals_the_analytic_x_mode():
    """A Y^2 table sampled from Y0 R0/(R0 + x): mode XT reproduces mode X (end states, optical path) to the
    accuracy of the spline; and the poloidal field of a real equilibrium shifts the right-hand cut-off."""
    xs = np.linspace(-10, 10, 401)
    d = xs[1] - xs[0]
    X = 0.3 * np.exp(-(xs[None, :] ** 2 + xs[:, None] ** 2) / 14.0)
    Y0, R0 = 0.45, 55.0
    Y2 = ((Y0 * R0 / (R0 + xs)) ** 2)[None, :] + 0 * xs[:, None]
    a = rk.Medium(X, xs[0], xs[0], d, d, mode="X", Y0=Y0, R0=R0)
    b = rk.Medium(X, xs[0], xs[0], d, d, mode="XT", Y2=Y2)
    rng = np.random.default_rng(2)
    n = 200
    x0, y0, th = rng.uniform(-6, 6, n), np.full(n, -9.5), rng.uniform(-0.4, 0.4, n)
    N = np.array([np.sqrt(a.eval(x0[i], y0[i])[3]) for i in range(n)])
    # end the rays ON a receiver line: the free end state is wherever the last step happened to land
    ra = rk.trace(a, x0, y0, N * np.sin(th), N * np.cos(th), rtol=1e-10, threads=4, stop=(9.0, "y"))
    rb = rk.trace(b, x0, y0, N * np.sin(th), N * np.cos(th), rtol=1e-10, threads=4, stop=(9.0, "y"))
    assert np.array_equal(ra["status"], rb["status"])
    on = ra["status"] == rk.STOP_LINE
    assert on.sum() > n // 2
    assert np.max(np.abs(ra["end_state"] - rb["end_state"])[:, on]) < 1e-6
    assert np.max(np.abs(ra["phi"] - rb["phi"])[on]) < 1e-6
