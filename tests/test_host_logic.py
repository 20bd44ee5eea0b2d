"""Argument validation and host-side marshalling of the Python surface; nothing
here reaches a CUDA call (the checks run first)."""
import numpy as np
import pytest

from rayrtrace_b200 import sidtrace


def _arrs(L=8, dtype=np.float64):
    return [np.zeros(L, dtype) for _ in range(4)]


def test_trace_rejects_wrong_dtype_and_layout():
    ax = np.zeros((4, 6)); ay = np.zeros((4, 6))
    x, y, kx, ky = _arrs()
    with pytest.raises(TypeError):
        sidtrace.trace(x.astype(np.float32), y, kx, ky, 0.1, 1, 1, 7, 5, ax, ay)
    with pytest.raises(TypeError):
        sidtrace.trace(list(x), y, kx, ky, 0.1, 1, 1, 7, 5, ax, ay)
    with pytest.raises(ValueError):
        sidtrace.trace(np.zeros(16)[::2], y, kx, ky, 0.1, 1, 1, 7, 5, ax, ay)
    with pytest.raises(ValueError):
        sidtrace.trace(x, np.zeros(9), kx, ky, 0.1, 1, 1, 7, 5, ax, ay)
    with pytest.raises(ValueError):
        sidtrace.trace(x, y, kx, ky, 0.1, 1, 1, 8, 5, ax, ay)          # ax is not (ny-1)*(nx-1)
    with pytest.raises(TypeError):
        sidtrace.trace(x, y, kx, ky, 0.1, 1, 1, 7, 5, ax.astype(np.float32), ay)


def test_check_bounds_and_type():
    a = np.arange(35.0).reshape(5, 7)
    with pytest.raises(IndexError):
        sidtrace.check(7, 5, 5, 0, a)
    with pytest.raises(TypeError):
        sidtrace.check(7, 5, 0, 0, a.astype(np.int64))


def test_check_value_is_reference_layout():
    # a[i*nx + j] (sidtracemodule.c:66): a host-memory probe, runs without a GPU
    a = np.arange(35.0).reshape(5, 7)
    assert sidtrace.check(7, 5, 2, 3, a) == 17.0


def test_dropin_module_resolves():
    import importlib
    import sys
    import rayrtrace_b200
    sys.path.insert(0, rayrtrace_b200.dropin_path())
    try:
        sys.modules.pop("sidtrace", None)
        m = importlib.import_module("sidtrace")
        assert m.trace is sidtrace.trace and m.check is sidtrace.check
    finally:
        sys.path.remove(rayrtrace_b200.dropin_path())
        sys.modules.pop("sidtrace", None)


def test_vacuum_k_rules():
    """ray.py:34-51: directions are normalised to w/c; |k| > 2 at the default freq infers the frequency."""
    from rayrtrace_b200.ray import _vacuum_k
    c = 299792458.0 * 100
    kx, ky, f = _vacuum_k(0.0, 2.0, 136e9)
    assert kx == 0.0 and ky == pytest.approx(2 * np.pi * 136e9 / c, rel=1e-15) and f == 136e9
    kx, ky, f = _vacuum_k(3.0, 4.0, 135e9)
    assert (kx, ky) == (3.0, 4.0) and f == pytest.approx(5.0 * c / 2 / np.pi, rel=1e-15)


def test_rk_medium_argument_validation_needs_no_device():
    """physics.Medium rejects bad tables before any CUDA call (shape, grid uniformity, mode, interpolant)."""
    from rayrtrace_b200 import physics
    x = np.linspace(-1, 1, 9)
    X = np.zeros((9, 9))
    with pytest.raises(ValueError, match="shape"):
        physics.Medium(np.zeros((9, 8)), x, x)
    with pytest.raises(ValueError, match="uniform"):
        physics.Medium(X, x ** 3, x)
    with pytest.raises(ValueError, match="increasing"):
        physics.Medium(X, x[::-1], x)
    with pytest.raises(ValueError, match="mode"):
        physics.Medium(X, x, x, mode="Z")
    with pytest.raises(ValueError, match="interp"):
        physics.Medium(X, x, x, interp="quintic")
    with pytest.raises(ValueError, match="alpha"):
        physics.Medium(X, x, x, alpha=np.zeros((3, 3)))
    with pytest.raises(ValueError, match="2 x 2"):
        physics.Medium(np.zeros((1, 9)), x, x[:1])
    # the tabulated-field X mode (STB_RK_MODE_X_TABLE)
    eq = physics.Equilibrium.solovev(nr=9, nz=9, nf=5)
    with pytest.raises(ValueError):
        physics.Medium(X, x, x, mode="O", Y2=np.zeros_like(X))           # a field table belongs to the X mode
    with pytest.raises(ValueError):
        physics.Medium(X, x, x, mode="X", Y2=np.zeros_like(X), equilibrium=eq, freq=60e9)
    with pytest.raises(ValueError):
        physics.Medium(X, x, x, mode="X", equilibrium=eq)                # Y = f_ce / freq needs the frequency
    with pytest.raises(ValueError):
        physics.Medium(X, x, x, mode="X", Y2=np.zeros((3, 3)))


def test_solovev_equilibrium_tables():
    """physics.Equilibrium.solovev: psi = 0 on the axis, ~psi_b on the midplane at R0 + a, and the analytic field is
    -psi_Z/R, psi_R/R of the table (centred differences) with F = B0 R0."""
    from rayrtrace_b200 import physics
    eq = physics.Equilibrium.solovev(R0=0.55, a=0.08, kappa=1.3, psi_b=2e-3, nr=201, nz=201)
    R = np.linspace(eq.rleft, eq.rleft + eq.rdim, eq.nx)
    Z = np.linspace(eq.zmid - eq.zdim / 2, eq.zmid + eq.zdim / 2, eq.ny)
    i0, j0 = np.argmin(abs(R - 0.55)), np.argmin(abs(Z))
    assert abs(eq.psi[i0, j0]) < 1e-12 and eq.psi.min() >= 0
    assert abs(eq.psi[np.argmin(abs(R - 0.63)), j0] / 2e-3 - 1) < 0.2
    dR, dZ = R[1] - R[0], Z[1] - Z[0]
    pR = (eq.psi[2:, 1:-1] - eq.psi[:-2, 1:-1]) / (2 * dR)
    pZ = (eq.psi[1:-1, 2:] - eq.psi[1:-1, :-2]) / (2 * dZ)
    RR, ZZ = np.meshgrid(R[1:-1], Z[1:-1], indexing="ij")
    BR, BZ, Bp = eq.analytic_B(RR, ZZ)
    assert np.max(np.abs(BR + pZ / RR)) < 1e-3 * np.max(np.abs(BR))
    assert np.max(np.abs(BZ - pR / RR)) < 1e-3 * np.max(np.abs(BZ))
    assert np.allclose(Bp * RR, 2.2 * 0.55)


def test_coherent_order_groups_neighbouring_rays():
    """physics.coherent_order (host logic, plain torch): a permutation; rays that share launch cell and
    direction end up next to each other; NaN / out-of-table launches do not break it."""
    import types
    import torch
    from rayrtrace_b200 import physics
    med = types.SimpleNamespace(xorg=-10.0, yorg=-10.0, dx=0.02, dy=0.02, nx=1001, ny=1001)
    g = torch.Generator().manual_seed(0)
    n = 20000
    x0 = (torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * 10
    y0 = torch.full_like(x0, -9.5)
    th = (torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * 0.6
    x0[3] = float("nan"); x0[4] = 1e30; y0[5] = -1e30
    perm = physics.coherent_order(med, x0, y0, torch.sin(th), torch.cos(th))
    assert sorted(perm.tolist()) == list(range(n))
    keep = torch.isfinite(x0[perm]) & (x0[perm].abs() < 100) & (y0[perm].abs() < 100)
    xs, ts = x0[perm][keep], th[perm][keep]
    jump = ((xs[1:] - xs[:-1]).abs() / 10 + (ts[1:] - ts[:-1]).abs() / 0.6)
    rnd = ((x0[1:] - x0[:-1]).abs() / 10 + (th[1:] - th[:-1]).abs() / 0.6)[torch.isfinite(x0[1:]) & torch.isfinite(x0[:-1])]
    assert jump.median() < 0.05 * rnd.median()



def test_pick_cells_follows_the_launch_directions():
    """sidtrace.pick_cells (host logic, no device): the cell-table layout follows the rays -- y-major when the
    launch wave vectors point mostly along y, the reference's x-major rows (sidtracemodule.c:38) otherwise."""
    from rayrtrace_b200 import sidtrace
    rng = np.random.default_rng(3)
    ang = rng.uniform(-0.4, 0.4, 100_000)
    assert sidtrace.pick_cells(np.sin(ang), np.cos(ang)) == "y"          # a fan shooting upwards (the FT-2 antennas)
    assert sidtrace.pick_cells(np.cos(ang), np.sin(ang)) == "x"
    assert sidtrace.pick_cells(-np.sin(ang), -np.cos(ang)) == "y"        # downwards counts the same
    assert sidtrace.pick_cells(np.empty(0), np.empty(0)) == "x"
    assert sidtrace.pick_cells(np.array([1.0]), np.array([1.0])) == "x"  # a tie keeps the reference layout


def test_coherent_order_of_the_parity_path_groups_launch_cells():
    """sidtrace.coherent_order (host logic): a permutation that puts rays of the same 4 x 4-cell block and direction
    sector next to each other, whatever NaN / out-of-grid launches it is given."""
    from rayrtrace_b200 import sidtrace
    rng = np.random.default_rng(9)
    n = 20000
    x0, y0 = rng.uniform(0, 200, n), rng.uniform(0, 200, n)
    ang = rng.uniform(-np.pi, np.pi, n)
    x0[:3] = [np.nan, -5.0, 1e9]
    p = sidtrace.coherent_order(x0, y0, np.sin(ang), np.cos(ang))
    assert sorted(p.tolist()) == list(range(n))
    bx, by = (np.nan_to_num(x0[p]).clip(0, 65535).astype(int) >> 2), (np.nan_to_num(y0[p]).clip(0, 65535).astype(int) >> 2)
    same_block = (bx[1:] == bx[:-1]) & (by[1:] == by[:-1])
    assert same_block.mean() > 0.8                       # 2500 blocks, 20000 rays: neighbours mostly share a block
    jump = np.abs(np.diff(bx)) + np.abs(np.diff(by))
    assert np.median(jump[~same_block]) <= 2             # and the curve moves to an adjacent block
