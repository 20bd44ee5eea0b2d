"""Synthetic workloads shaped like BASELINE.json's configs (there are no public
data sets for this path): an FT-2-like shaped density on a square grid and
the seven-antenna Gaussian-beam fan of ft2.phases (ft2.py:12-16, 47-54) sampled
with as many rays as asked.

Everything here is host-side input generation; nothing is traced.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy import constants as const

from .nfield import Nfield
from .launch import launch_states

R_DIA = 7.85      # limiter radius, cm (ft2.py:6)
R_ANT = 10.0      # antenna radius, cm (ft2.py:7)
BEAM_FREQ = 136.268e9   # beam.py:30
WAIST = 0.35      # beam.py:31
FOCUS = -1.0      # beam.py:31: the waist lies 1 cm behind the antenna
DIST = 3.0        # beam.py:25: rays start on the wavefront 3 cm from the focus
AMPBND = 1e-3     # beam.py:26


def shaped_density(nx, *, ny=None, n0=4e13, edge=1e12, shift=0.6, elon=1.1, alpha=1.0, half=10.0):
    """Parabolic profile on Shafranov-shifted, elongated flux surfaces:
    s = rho^2 solves (x - shift*(1 - s))^2 + (y/elon)^2 = R_DIA^2 * s (closed form).
    Returns (n[ny, nx], x, y) in cm^-3 / cm.  With shift=0, elon=1 it is the
    parabolic FT-2 profile of SURVEY 8d config 1."""
    ny = nx if ny is None else ny
    x = np.linspace(-half, half, nx)
    y = np.linspace(-half, half, ny)
    X, Y = np.meshgrid(x, y)
    C = (X - shift) ** 2 + (Y / elon) ** 2
    if shift == 0:
        s = C / R_DIA ** 2
    else:
        A = shift * shift
        B = 2 * shift * (X - shift) - R_DIA ** 2          # < 0 on the whole grid
        disc = np.maximum(B * B - 4 * A * C, 0.0)
        s = 2 * C / (-B + np.sqrt(disc))                  # smaller root, cancellation-free
    n = n0 * np.maximum(0.0, 1 - s) ** alpha + edge
    return n, x, y


def antenna_positions():
    """The seven FT-2 antennas on the lower half of the r = 10 cm circle (ft2.py:12-16, 30-31)."""
    xa = np.linspace(-6.9, 6.9, 7)
    return xa, -np.sqrt(R_ANT ** 2 - xa ** 2)


def gaussian_fan(nrays, *, freq=BEAM_FREQ, n_edge=1.0, index=None):
    """Start points, unit normals and amplitudes of `nrays` rays spread over the seven
    upward-pointing beams, equidistant in arc length along the (spherical) wavefront at
    DIST from the focus -- the analytic stand-in for Beam.rays (beam.py:278-331).
    Returns (x, y, kx, ky, amp) in lab coordinates; `index` (global ray indices) restricts
    the result to a shard without materialising the whole fan."""
    lam = const.speed_of_light * 100 / freq
    y_r = np.pi * WAIST ** 2 * n_edge / lam                # beam.py:134
    wa = WAIST * np.sqrt(1 + DIST ** 2 / y_r ** 2)         # beam.py:183
    rc = DIST * (1 + y_r ** 2 / DIST ** 2)                 # beam.py:185 curvature radius
    xmax = np.sqrt(-wa ** 2 * np.log(AMPBND))              # beam.py:218
    th_max = np.arcsin(min(xmax / rc, 0.999))
    xa, ya = antenna_positions()
    per = np.full(7, nrays // 7)
    per[: nrays % 7] += 1
    first = np.concatenate(([0], np.cumsum(per)))          # ray index where each antenna's fan starts
    idx = np.arange(nrays, dtype=np.int64) if index is None else np.asarray(index, np.int64)
    ant = np.searchsorted(first, idx, side="right") - 1
    loc = idx - first[ant]
    cnt = per[ant]
    # np.linspace(-th_max, th_max, cnt)[loc], element for element
    step = np.where(cnt > 1, (2 * th_max) / np.maximum(cnt - 1, 1), 0.0)
    th = np.where(cnt > 1, loc * step + (-th_max), 0.0)
    th = np.where((cnt > 1) & (loc == cnt - 1), th_max, th)
    xb = rc * np.sin(th)
    yb = DIST - rc + rc * np.cos(th)                       # along-axis distance from the focus
    x = xa[ant] + xb
    y = ya[ant] + (yb + FOCUS)                             # beam2lab with k0 = (0, 1) (beam.py:117-131)
    amp = np.sqrt(WAIST / wa) * np.exp(-xb ** 2 / wa ** 2)  # beam.py:188
    return x, y, np.sin(th), np.cos(th), amp


@dataclass
class BeamCase:
    nfield: Nfield
    launch: tuple           # (xg0, yg0, kx0, ky0) float64 arrays, kernel units
    amp: np.ndarray
    x0: np.ndarray          # physical launch points
    y0: np.ndarray
    nrays: int
    L: int
    st: float
    dx: float
    dy: float
    freq: float
    name: str


def default_cap(nx: int) -> int:
    """States per ray for the synthetic beams: 1.5 * nx rounded up to 128.  A vacuum ray
    crosses the grid in ~nx steps (SURVEY 3.3 i); rays that need more hit the cap exactly
    like the reference does at Ray.MaxStep."""
    return int(-(-(3 * nx // 2) // 128) * 128)


def beam_case(nx=1001, nrays=1_000_000, *, seed=0, n0=4e13, shift=0.6, elon=1.1, L=None, name=None,
              index=None, device_launch=False) -> BeamCase:
    """BASELINE configs 3/5 shape: shaped density on an nx^2 grid + seven-beam fan of `nrays` rays.
    `seed` jitters the start angle of every ray by a fraction of the ray spacing so no two
    rays are identical.  `index` (global ray indices) builds only that shard of the fan (the jitter
    is a hash of the global index then, so shards of different ranks are consistent);
    `device_launch` prepares the launch states on the GPU (`launch_states_device`)."""
    n, x, y = shaped_density(nx, n0=n0, shift=shift, elon=elon)
    nf = Nfield(n, x=x, y=y)
    xs, ys, kx, ky, amp = gaussian_fan(nrays, index=index)
    if index is None:
        jit = np.random.default_rng(seed).uniform(-1e-4, 1e-4, nrays)
    else:   # counter-based jitter: a function of (seed, global ray index) only
        h = (np.asarray(index, np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(31)
        jit = ((h >> np.uint64(11)).astype(np.float64) / 2.0 ** 53 * 2 - 1) * 1e-4
    kx, ky = kx * np.cos(jit) - ky * np.sin(jit), kx * np.sin(jit) + ky * np.cos(jit)
    if device_launch:
        from .ray import launch_states_device
        xg, yg, kxs, kys, freq = launch_states_device(nf, xs, ys, kx, ky, BEAM_FREQ)
        xg, yg, kxs, kys = (t.cpu().numpy() for t in (xg, yg, kxs, kys))
    else:
        xg, yg, kxs, kys, freq = launch_states(nf, xs, ys, kx, ky, BEAM_FREQ)
    nrays = len(xs)
    return BeamCase(nfield=nf, launch=(xg, yg, kxs, kys), amp=amp, x0=xs, y0=ys, nrays=nrays,
                    L=L or default_cap(nx), st=float(nf.st), dx=float(nf.dx[0]), dy=float(nf.dy[0]), freq=freq,
                    name=name or f"ft2-7beam-fan-{nrays}rays-{nx}x{nx}")


def sweep_case(nx=201, nrays=100_000, *, seed=0, L=None, device_launch=False) -> BeamCase:
    """BASELINE config 4 shape: launch-angle / density-regime sweep with cutoff-heavy
    trajectories: an OVER-dense core (n0 = 1.3 n_c) reflects head-on rays, grazing rays
    skim it, and launch angles span +-0.3 rad (Beam.k0 rotation, beam.py:96-100)."""
    n, x, y = shaped_density(nx, n0=3.0e14, shift=0.3, elon=1.05)
    nf = Nfield(n, x=x, y=y)
    rng = np.random.default_rng(seed)
    xs = rng.uniform(-7.5, 7.5, nrays)
    ys = np.full(nrays, -9.5)
    ang = rng.uniform(-0.3, 0.3, nrays)
    if device_launch:
        from .ray import launch_states_device
        xg, yg, kxs, kys, freq = launch_states_device(nf, xs, ys, np.sin(ang), np.cos(ang), BEAM_FREQ)
        xg, yg, kxs, kys = (t.cpu().numpy() for t in (xg, yg, kxs, kys))
    else:
        xg, yg, kxs, kys, freq = launch_states(nf, xs, ys, np.sin(ang), np.cos(ang), BEAM_FREQ)
    return BeamCase(nfield=nf, launch=(xg, yg, kxs, kys), amp=np.ones(nrays), x0=xs, y0=ys, nrays=nrays,
                    L=L or 4 * default_cap(nx), st=float(nf.st), dx=float(nf.dx[0]), dy=float(nf.dy[0]), freq=freq,
                    name=f"cutoff-sweep-{nrays}rays-{nx}x{nx}")


def trapped_case(nx=257, nrays=500_000, *, seed=0, L=8192, trapped_frac=0.03) -> BeamCase:
    """Heavy-tailed step counts (the stand-in for north_star's "resonance-heavy" rays):
    an over-dense ring around a hollow core is a closed cavity.  Most rays are launched outside
    and leave after ~nx steps; a few per cent start inside the cavity and bounce between its
    walls for thousands of steps (until the scheme's numerical heating lets them over the
    barrier, ~15-20x the typical ray)."""
    half = 10.0
    x = np.linspace(-half, half, nx)
    y = np.linspace(-half, half, nx)
    X, Y = np.meshgrid(x, y)
    r = np.sqrt(X ** 2 + Y ** 2)
    n = 4.0e14 * np.exp(-((r - 6.0) / 2.5) ** 2) + 1e12     # over-dense ring (n_c = 2.3e14), hollow core: a closed cavity
    nf = Nfield(n, x=x, y=y)
    rng = np.random.default_rng(seed)
    ntr = int(nrays * trapped_frac)
    xs = rng.uniform(-7.5, 7.5, nrays)
    ys = np.full(nrays, -9.5)
    ang = rng.uniform(-0.3, 0.3, nrays)
    # trapped population: start near the axis of the hollow, any direction
    idx = rng.choice(nrays, ntr, replace=False)
    rr = rng.uniform(0.0, 2.5, ntr)
    th = rng.uniform(0, 2 * np.pi, ntr)
    xs[idx], ys[idx] = rr * np.cos(th), rr * np.sin(th)
    ang[idx] = rng.uniform(0, 2 * np.pi, ntr)
    xg, yg, kxs, kys, freq = launch_states(nf, xs, ys, np.sin(ang), np.cos(ang), BEAM_FREQ)
    return BeamCase(nfield=nf, launch=(xg, yg, kxs, kys), amp=np.ones(nrays), x0=xs, y0=ys, nrays=nrays,
                    L=L, st=float(nf.st), dx=float(nf.dx[0]), dy=float(nf.dy[0]), freq=freq,
                    name=f"trapped-{nrays}rays-{nx}x{nx}")
