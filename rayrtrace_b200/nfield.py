"""Host-side field producer: the reference's ``den.Nfield`` restated.

It defines the read-only inputs of the hot path -- ``ax``, ``ay`` (acceleration
per cell) and the global step ``st`` -- so its arithmetic follows the
reference operation for operation (den.py:58-66 forward differences,
den.py:84-107 scaling and step, den.py:335-348 critical density); the results
are bit-identical to the reference's (tests/test_nfield.py pins them against
fixtures produced by the unmodified den.py).

This is O(grid) work done once per (density, frequency): it stays on the host,
like in the reference.  ``Nfield.device_field()`` uploads ``ax, ay`` once and
returns the HBM-resident ``sidtrace.Field`` the kernels read.
"""
from __future__ import annotations

import numpy as np
from scipy import constants as const
from scipy.interpolate import RectBivariateSpline, interp1d

# CGS electron charge as the reference hard-codes it (den.py:340)
_E_CGS = 4.8032e-10
FILL_VALUE = 1e12     # den.py:71: density the reference's interpolant returns outside the grid


def critical_den(freq=136.29e9, den_unit="cm^-3"):
    """n_c = m_e w^2 / (4 pi e^2) in CGS (den.py:335-348); same operation order."""
    w = 2 * np.pi * freq
    scale = 1
    if den_unit == "m^-3":
        scale = 1e6
    return scale * const.electron_mass * 1000 * w ** 2 / (4 * np.pi * _E_CGS ** 2)


class Nfield:
    """Density grid -> acceleration field + step.

    Same constructor and attributes as the reference class (den.py:7-107):
    ``n, x, y, nx, ny, dx, dy, dn_x, dn_y, x2grid, y2grid, n_interpolant,
    min_dxy, freq, ax, ay, st`` and methods ``refractive_index``,
    ``prep_raytrace``.  Unlike the reference (which returns a half-built
    object on a shape mismatch, den.py:38-40) bad input raises ValueError.
    """

    def __init__(self, den, *, nx=100, ny=100, x=None, y=None,
                 freq=136.26e9, len_unit="cm", den_unit="cm^-3"):
        self.len_unit = len_unit
        self.den_unit = den_unit
        if isinstance(den, dict):
            # den.py:17-27
            self.n = np.array(den["n"], np.float64)
            self.x = np.array(den["x"])
            self.y = np.array(den["y"])
            self.nx, self.ny = len(self.x), len(self.y)
        elif isinstance(den, np.ndarray) and den.ndim == 2:
            # den.py:28-52: rows are y
            self.n = np.array(den, np.float64)
            self.ny, self.nx = den.shape
            self.x = self._axis(x, self.nx, den.size, "x")
            self.y = self._axis(y, self.ny, den.size, "y")
        else:
            raise ValueError("den must be a 2-D ndarray or a dict with keys n, x, y")
        if self.n.shape != (self.ny, self.nx):
            raise ValueError(f"density shape {self.n.shape} != (ny, nx) = ({self.ny}, {self.nx})")
        if self.nx < 2 or self.ny < 2:
            raise ValueError("the grid needs at least 2 nodes per axis")

        # grid step and forward-difference gradients on the lower-left corner (den.py:59-66)
        self.dx = self.x[1:] - self.x[:-1]
        self.dy = self.y[1:] - self.y[:-1]
        dxx, dyy = np.meshgrid(self.dx, self.dy)
        self.dn_x = (self.n[:-1, 1:] - self.n[:-1, :-1]) / dxx
        self.dn_y = (self.n[1:, :-1] - self.n[:-1, :-1]) / dyy

        # physical -> grid coordinates (den.py:69-70)
        self.x2grid = interp1d(self.x, np.arange(0, self.nx), fill_value="extrapolate")
        self.y2grid = interp1d(self.y, np.arange(0, self.ny), fill_value="extrapolate")
        # den.py:71 used interp2d(kind='cubic'), removed in SciPy 1.14; the cubic
        # RectBivariateSpline is its documented replacement.  Only the launch
        # refractive index (ray.py:62) reads it.
        self._spline = None

        self.min_dxy = min(self.dx.min(), self.dy.min())
        self._device_field = None
        self.prep_raytrace(freq=freq)

    @staticmethod
    def _axis(v, n, size, name):
        if v is None:
            return np.arange(n)
        v = np.atleast_1d(v)
        if len(v) != n and v.size != size:
            raise ValueError(f"{name} has {len(v)} nodes, density has {n}")
        return v

    def n_interpolant(self, x, y):
        """Cubic interpolation of n at (x, y); scalar queries return shape (1,)
        like the reference's interp2d object (den.py:71)."""
        if self._spline is None:
            self._spline = RectBivariateSpline(np.asarray(self.x, float), np.asarray(self.y, float),
                                               self.n.T, kx=3, ky=3, s=0)
        xa = np.atleast_1d(np.asarray(x, float))
        ya = np.atleast_1d(np.asarray(y, float))
        out = self._spline(xa, ya).T
        # den.py:71: interp2d(..., fill_value=1e12) -- outside the grid the reference reads 1e12
        # (RectBivariateSpline would clamp to the boundary)
        xo = (xa < self.x.min()) | (xa > self.x.max())
        yo = (ya < self.y.min()) | (ya > self.y.max())
        if xo.any() or yo.any():
            out = np.where(yo[:, None] | xo[None, :], FILL_VALUE, out)
        if out.shape[0] == 1:
            return out[0]
        if out.shape[1] == 1:
            return out[:, 0]
        return out

    def n_interpolant_points(self, x, y):
        """Cubic interpolation of n at the point pairs (x[i], y[i])."""
        self.n_interpolant(self.x[0], self.y[0])   # builds the spline
        x, y = np.asarray(x, float), np.asarray(y, float)
        out = self._spline.ev(x, y)
        outside = (x < self.x.min()) | (x > self.x.max()) | (y < self.y.min()) | (y > self.y.max())
        return np.where(outside, FILL_VALUE, out) if np.any(outside) else out

    def refractive_index(self, freq=136.26e9, x=None, y=None):
        """N = sqrt(1 - n/n_c) (den.py:76-82); boundary density when x or y is None."""
        if x is None or y is None:
            den = self.n[0, 0]
        else:
            den = self.n_interpolant(x, y)
        nc = critical_den(freq, den_unit=self.den_unit)
        return np.sqrt(1 - den / nc)

    def prep_raytrace(self, freq=136.26e9):
        """ax, ay and the global step st (den.py:84-107), same operation order."""
        self.freq = freq
        w = 2 * np.pi * self.freq
        c = const.speed_of_light * 100
        nc = critical_den(freq, den_unit=self.den_unit)

        self.ax = -self.dn_x / 2 / nc * (w / c) ** 2
        self.ay = -self.dn_y / 2 / nc * (w / c) ** 2

        a_max = max(np.max(abs(self.ax)), np.max(abs(self.ay)))
        b_max = w / c
        if a_max == 0:
            st_grad = np.inf
        else:
            # positive root of a_max*st^2 + b_max*st = min_dxy
            st_grad = (-b_max + np.sqrt(b_max ** 2 + 4 * a_max * self.min_dxy)) / 2 / a_max
        self.st = min(st_grad, self.min_dxy * c / w)
        self._device_field = None

    def close(self):
        """Release the HBM copy of the field (it is re-uploaded on next use)."""
        if self._device_field is not None:
            self._device_field.close()
            self._device_field = None
        self._device_launch = None

    def device_launch_tables(self):
        """CUDA copies of what the device launcher needs: node coordinates and the bicubic spline of
        n (knots, coefficients) fitted once on the host."""
        tabs = getattr(self, "_device_launch", None)
        if tabs is None:
            import torch
            self.n_interpolant(self.x[0], self.y[0])     # builds the spline
            tx, ty, c = self._spline.tck
            dev = torch.device("cuda", torch.cuda.current_device())
            tabs = {k: torch.from_numpy(np.ascontiguousarray(v, np.float64)).to(dev)
                    for k, v in (("x", self.x), ("y", self.y), ("tx", tx), ("ty", ty), ("c", c))}
            self._device_launch = tabs
        return tabs

    # ---- B200 side -------------------------------------------------------
    def device_field(self):
        """The HBM-resident copy of (ax, ay) the kernels gather from (uploaded once)."""
        if self._device_field is None:
            from .sidtrace import Field
            self._device_field = Field(np.ascontiguousarray(self.ax), np.ascontiguousarray(self.ay),
                                       self.nx, self.ny)
        return self._device_field
