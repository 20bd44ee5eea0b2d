"""Shared helpers for the tests: synthetic densities / fields / launch sets."""
import hashlib

import numpy as np

KAT_NAMES = ("G1", "G2", "G3", "G4", "G5", "G6a", "G6b", "G6c")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ft2_density(nx, n0, edge=1e12, r_dia=7.85, half=10.0):
    """Same analytic profile tests/golden/make_golden.py used."""
    x = np.linspace(-half, half, nx)
    y = np.linspace(-half, half, nx)
    X, Y = np.meshgrid(x, y)
    n = n0 * np.maximum(0, 1 - (X ** 2 + Y ** 2) / r_dia ** 2) + edge
    return n, x, y


def random_field(rng, nx, ny, scale=0.3, smooth=True):
    """A random acceleration field of shape (ny-1, nx-1) with O(scale) values."""
    ax = rng.normal(0, scale, (ny - 1, nx - 1))
    ay = rng.normal(0, scale, (ny - 1, nx - 1))
    if smooth:
        for a in (ax, ay):
            a[:] = (a + np.roll(a, 1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 0) + np.roll(a, -1, 1)) / 5
    return np.ascontiguousarray(ax), np.ascontiguousarray(ay)


def random_launch(rng, n, nx, ny, kmag=1.0):
    x0 = rng.uniform(0, nx - 1, n)
    y0 = rng.uniform(0, ny - 1, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    k = kmag * rng.uniform(0.2, 1.5, n)
    return x0, y0, k * np.cos(ang), k * np.sin(ang)


def traj_digest(traj, nsteps):
    """SHA-256 per ray of states 0..nsteps (as make_golden.py computed them)."""
    return np.array([sha(traj[r, :int(nsteps[r]) + 1]) for r in range(traj.shape[0])])
