"""Oracle for the receiver-line crossing: what the reference's
``Ray.addpoint(val, axis)`` (ray.py:131-224) reports for the FIRST crossing of
the line ``y = val`` (or ``x = val``), as a pure function of a kept trajectory.

TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/ft2_cross.npz, produced by
the unmodified reference ``Ray.addpoint`` (tests/golden/make_golden.py).

Restated, not inserted: the reference inserts the crossing point into the ray's
arrays and returns its index; ``Beam.field_r`` (beam.py:359-369) then reads
``r.x[ind]`` and ``r.ph[ind]``.  This function returns those values directly.

Two documented deviations from the reference text:
  * the cell of the segment start is ``trunc`` of the grid coordinate the tracer
    itself used; the reference maps the physical point back through
    ``x2grid = interp1d(...)`` (ray.py:166-167), identical unless the state sits
    within an ulp of a cell edge;
  * squares are products; the reference's ``ky ** 2`` on NumPy scalars goes
    through libm ``pow`` (ray.py:177-178), which can differ by one ulp.
Hence the tolerance for this path is 1e-9 relative (north_star), not bitwise.
"""
from __future__ import annotations

import math

import numpy as np


def _solve(c0, k, a, val, st1):
    """ray.py:174-193: smallest non-negative root of k*s + a*s^2/2 = val - c0."""
    st = st1
    if a != 0:
        disc = k * k + 2 * a * (val - c0)
        if disc < 0:
            return math.nan                      # reference: empty .min() -> ValueError
        sq = math.sqrt(disc)
        roots = [r for r in ((-k - sq) / a, (-k + sq) / a) if r >= 0]
        if not roots:
            return math.nan
        st = min(roots)
        if st == 0 and val != c0 and k != 0:
            st = (val - c0) / k
    elif k != 0:
        st = (val - c0) / k
    return st


def crossing(traj_kept, x0p, y0p, dx, dy, st, ax, ay, val, axis="y"):
    """traj_kept: [n, 4] kept states in grid units.  Returns dict(count, index, exact,
    extrapolated, x, y, kx, ky, lax, ph); count == 0 -> the reference returns nan."""
    t = np.asarray(traj_kept, np.float64)
    n = t.shape[0]
    out = dict(count=0, index=-1, exact=0, extrapolated=0, x=math.nan, y=math.nan, kx=math.nan, ky=math.nan,
               lax=math.nan, ph=math.nan)
    if n == 0:
        return out
    xs = x0p + (t[:, 0] - t[0, 0]) * dx          # ray.py:105-106
    ys = y0p + (t[:, 1] - t[0, 1]) * dy
    kx, ky = t[:, 2], t[:, 3]
    k = np.sqrt(kx * kx + ky * ky)
    kavr = (k[:-1] + k[1:]) / 2
    lax = np.cumsum(np.insert(st * kavr, 0, 0))  # ray.py:117-120 with st[:-1] == st
    ph = np.cumsum(np.insert(np.diff(lax) * kavr, 0, 0))
    c = ys if axis == "y" else xs
    kc = ky if axis == "y" else kx
    eq = np.where(c == val)[0]
    if len(eq):                                   # ray.py:142-143
        i = int(eq[0])
        out.update(count=len(eq), index=i, exact=1, x=xs[i], y=ys[i], kx=kx[i], ky=ky[i], lax=lax[i], ph=ph[i])
        return out
    if np.any(c < val) and np.any(val < c):       # ray.py:144-146
        i_n = np.where((c - val)[:-1] * (c - val)[1:] < 0)[0]
        st1 = st
    elif (val - c[-1]) * kc[-1] > 0:              # ray.py:147-148
        i_n = np.array([n - 1])
        st1 = 0.0
        out["extrapolated"] = 1
    else:
        return out
    if len(i_n) == 0:
        return out
    i = int(i_n[0])
    ix, iy = int(t[i, 0]), int(t[i, 1])           # trunc: states are >= 0 inside the grid
    a_x, a_y = float(ax[iy, ix]), float(ay[iy, ix])
    s = _solve(c[i], kc[i], a_y if axis == "y" else a_x, val, st1)
    if math.isnan(s) or s == 0:
        out["extrapolated"] = 0
        return out
    kxn = kx[i] + a_x * s                         # ray.py:212-216
    kyn = ky[i] + a_y * s
    xn = xs[i] + kx[i] * s + a_x / 2 * s * s
    yn = ys[i] + ky[i] * s + a_y / 2 * s * s
    kn = math.sqrt(kxn * kxn + kyn * kyn)
    kav = (k[i] + kn) / 2
    laxn = lax[i] + s * kav
    phn = ph[i] + (laxn - lax[i]) * kav
    out.update(count=len(i_n), index=i + 1, x=xn, y=yn, kx=kxn, ky=kyn, lax=laxn, ph=phn)
    return out


def ray_after_addpoint(traj_kept, x0p, y0p, dx, dy, st, ax, ay, val, axis="y"):
    """Arrays (x, y, ph) of the ray as the reference holds them after ``addpoint(val, axis)``
    inserted the FIRST crossing point (ray.py:198-216) and ``ph`` was recomputed from the modified
    ``st`` (ray.py:117-126).  Rays without a crossing (or with an exact hit) are returned unchanged."""
    t = np.asarray(traj_kept, np.float64)
    n = t.shape[0]
    xs = x0p + (t[:, 0] - t[0, 0]) * dx
    ys = y0p + (t[:, 1] - t[0, 1]) * dy
    kx, ky = t[:, 2].copy(), t[:, 3].copy()
    sts = np.full(n, float(st))
    sts[-1] = 0.0
    c = crossing(t, x0p, y0p, dx, dy, st, ax, ay, val, axis)
    if c["count"] and not c["exact"]:
        i = c["index"]                    # insertion index; segment start is i-1
        s_old = sts[i - 1]
        kav = (math.hypot(kx[i - 1], ky[i - 1]) + math.hypot(c["kx"], c["ky"])) / 2
        s = (c["lax"] - _lax_at(kx, ky, sts, i - 1)) / kav if kav else 0.0
        sts[i - 1] = s
        sts = np.insert(sts, i, s_old - s if i != n else 0.0)
        xs = np.insert(xs, i, c["x"]); ys = np.insert(ys, i, c["y"])
        kx = np.insert(kx, i, c["kx"]); ky = np.insert(ky, i, c["ky"])
    k = np.sqrt(kx * kx + ky * ky)
    kavr = (k[:-1] + k[1:]) / 2
    lax = np.cumsum(np.insert(sts[:-1] * kavr, 0, 0))
    ph = np.cumsum(np.insert(np.diff(lax) * kavr, 0, 0))
    return xs, ys, ph


def _lax_at(kx, ky, sts, j):
    k = np.sqrt(kx * kx + ky * ky)
    kavr = (k[:-1] + k[1:]) / 2
    return float(np.cumsum(np.insert(sts[:-1] * kavr, 0, 0))[j])


def interp_extrap(xn, yn, t):
    """scipy interp1d(kind='linear', fill_value='extrapolate') at one point (interp1d._call_linear,
    SciPy >= 1.12): node b is the first with t <= xn[b] (clipped to [1, n-1]);
    y = ((t - x_lo)/(x_hi - x_lo))*y_hi + ((x_hi - t)/(x_hi - x_lo))*y_lo."""
    b = int(np.clip(np.searchsorted(xn, t), 1, len(xn) - 1))
    w = xn[b] - xn[b - 1]
    return ((t - xn[b - 1]) / w) * yn[b] + ((xn[b] - t) / w) * yn[b - 1]


def phase_lookup(traj_kept, x0p, y0p, dx, dy, st, ax, ay, val, axis, targets):
    """Points of the (post-addpoint) ray at the target phases: [(x, y), ...] (beam.py:410-413)."""
    xs, ys, ph = ray_after_addpoint(traj_kept, x0p, y0p, dx, dy, st, ax, ay, val, axis)
    if len(ph) < 2:
        return [(math.nan, math.nan) for _ in targets]
    return [(interp_extrap(ph, xs, t), interp_extrap(ph, ys, t)) for t in targets]
