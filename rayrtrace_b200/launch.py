"""Launch-state preparation on the host -- pure NumPy / SciPy, no CUDA library needed.

What the reference's ``Ray.__init__`` does before it calls ``trace`` (ray.py:28-68): normalise the
direction to the vacuum wave number (or infer the frequency from |k|), scale by the local refractive
index, map the start point to grid units.  Kept apart from ``ray.py`` so that workload generators
(``synth``) and the CPU baseline arm of bench.py can build launch sets without loading
``libsidtrace_b200.so``.
"""
from __future__ import annotations

import numpy as np
from scipy import constants as const

_DEFAULT_FREQ = 135e9   # ray.py:15 -- also the sentinel of the "infer freq from |k|" rule (ray.py:34,47)


def _vacuum_k(kx0, ky0, freq):
    """ray.py:28-51: normalise the launch direction to w/c, or infer the
    frequency from |k| when |k| > 2 and freq is the default.  Returns
    (kx0, ky0, freq).  Scalar inputs use the same numpy calls as the reference."""
    c = const.speed_of_light * 100
    w = 2 * np.pi * freq
    if np.sqrt(kx0 ** 2 + ky0 ** 2) <= 2 or freq != _DEFAULT_FREQ:
        k = np.linalg.norm((kx0, ky0))
        kx0 = kx0 / k
        ky0 = ky0 / k
        kx0 = kx0 * w / c
        ky0 = ky0 * w / c
    elif freq == _DEFAULT_FREQ:
        w = np.sqrt(kx0 ** 2 + ky0 ** 2) * c
        freq = w / 2 / np.pi
    return kx0, ky0, freq


def launch_states(n_field, x0, y0, kx0, ky0, freq):
    """Vectorised ray.py:28-68: returns (xg0, yg0, kx, ky, freq) as the kernel
    expects them.  Direction vectors are normalised with sqrt(kx^2+ky^2)
    (np.linalg.norm per ray in the reference; <= 1 ulp apart)."""
    x0 = np.atleast_1d(np.asarray(x0, np.float64))
    y0 = np.atleast_1d(np.asarray(y0, np.float64))
    kx0 = np.broadcast_to(np.asarray(kx0, np.float64), x0.shape).copy()
    ky0 = np.broadcast_to(np.asarray(ky0, np.float64), x0.shape).copy()
    c = const.speed_of_light * 100
    w = 2 * np.pi * freq
    kn = np.sqrt(kx0 ** 2 + ky0 ** 2)
    if freq != _DEFAULT_FREQ or np.all(kn <= 2):
        kx = kx0 / kn * w / c
        ky = ky0 / kn * w / c
    else:
        raise ValueError("a batch must share one frequency: pass freq explicitly (ray.py:47-51 infers it per ray)")
    den = n_field.n_interpolant_points(x0, y0) if hasattr(n_field, "n_interpolant_points") else \
        np.array([np.asarray(n_field.n_interpolant(a, b)).reshape(-1)[0] for a, b in zip(x0, y0)])
    from .nfield import critical_den
    refr = np.sqrt(1 - den / critical_den(freq, den_unit=getattr(n_field, "den_unit", "cm^-3")))
    kx = kx * refr
    ky = ky * refr
    xg = np.ascontiguousarray(n_field.x2grid(x0), np.float64)
    yg = np.ascontiguousarray(n_field.y2grid(y0), np.float64)
    return xg, yg, np.ascontiguousarray(kx), np.ascontiguousarray(ky), freq
