"""Beam-level caller of the hot path: the reference's ``Beam`` (beam.py:22-444) and
``field_result`` (beam.py:447-472) with the serial ``for ...: Ray(...)`` loop of
beam.py:340-342 replaced by ONE batched launch (``trace_rays``).

SURVEY 8 f-2 / f-3 ("next" rows): everything here is O(rays) host work around the
O(rays x steps) GPU trace.  What differs from the reference, on purpose:

* The wavefront (beam.py:209-248).  The default ``Beam.front_method = "contour"`` follows the
  reference: the marching-squares iso-line of a 1000 x 1000 phase map
  (``skimage.measure.find_contours`` in the reference; skimage is not a dependency, so its
  published algorithm -- one vertex per straddled cell edge, linear interpolation -- is restated
  in ``_iso_line``) and a cubic interpolant through its ~1 200 vertices.  Only the ~6 000 map nodes
  next to the front are evaluated (``_iso_line_banded``): same vertices bit for bit, 3 ms instead
  of 0.1-0.2 s.  ``"newton"`` solves ``phase(x, y) = ph0`` per column on the analytic phase
  instead: the exact iso-phase line, 1e-6 cm away from the contour one -- but the contour's
  cell-periodic interpolation error tilts the launch normals by ~1e-4 rad and the interferometer
  phases by ~1e-3 rad, so parity with the reference needs the contour front.  The fixtures in
  tests/golden come from the reference's own finder.
* Rays live in HBM (``self.bundle``); ``self._rays`` materialises reference-shaped ``Ray``
  objects on demand.
"""
from __future__ import annotations

import numpy as np
from scipy import constants as const
from scipy.interpolate import interp1d
from scipy.optimize import curve_fit

from .ray import trace_rays


class Beam:
    _nfront = 1000   # beam.py:23  points of the wavefront finder
    _nray = 101      # beam.py:24  default number of rays
    _dist = 3        # beam.py:25  distance from the focus where rays start
    _ampbnd = 0.001  # beam.py:26  relative amplitude bounding the launch area
    _gouy = 1        # beam.py:27  Gouy phase switch
    front_method = "contour"  # the reference's marching-squares finder (beam.py:225-247); or "newton"
    _front_cache = {}         # wavefront interpolants by beam shape (see _wavefront)

    def __init__(self, x0=0, y0=0, kx0=0, ky0=1, freq=136.268e9, waist=0.35, focus=-1,
                 n_field=None, freq_unit="Hz", length_unit="cm"):
        self.x0, self.y0 = x0, y0
        self._kx0, self._ky0 = kx0, ky0
        self.freq = freq
        self.waist = waist
        self.focus = focus
        self.freq_unit = freq_unit
        self.length_unit = length_unit
        self.bundle = None
        self._ray_cache = None
        if n_field is not None:
            self.rays(rnum=Beam._nray, wave_front=None, dist=Beam._dist, ampbnd=Beam._ampbnd, n_field=n_field)

    # ---- units / direction (beam.py:47-100) ---------------------------------
    def _len_scale(self):
        return 1 if self.length_unit == "m" else 100

    @property
    def wave_len(self):
        return self._len_scale() * const.speed_of_light / self.freq

    @wave_len.setter
    def wave_len(self, lw):
        if 0 < lw < 0.1:
            self.freq = self._len_scale() * const.speed_of_light / lw

    @property
    def k0(self):
        return self._kx0, self._ky0

    @k0.setter
    def k0(self, k=(0, 1)):
        kn = np.linalg.norm(k)
        self._kx0, self._ky0 = k[0] / kn, k[1] / kn

    @property
    def kx0(self):
        return self._kx0

    @kx0.setter
    def kx0(self, kx):
        if abs(kx) >= 1:
            self._kx0, self._ky0 = kx / abs(kx), 0
        else:
            self._kx0 = kx
            self._ky0 = self._ky0 / abs(self._ky0) * np.sqrt(1 - kx ** 2)

    @property
    def ky0(self):
        return self._ky0

    @ky0.setter
    def ky0(self, ky):
        if abs(ky) >= 1:
            self._ky0, self._kx0 = ky / abs(ky), 0
        else:
            self._ky0 = ky
            self._kx0 = self._kx0 / abs(self._kx0) * np.sqrt(1 - ky ** 2)

    # ---- frames (beam.py:103-131) -------------------------------------------
    def lab2beam(self, xl, yl):
        xb = self._ky0 * (xl - self.x0) - self._kx0 * (yl - self.y0)
        yb = self._kx0 * (xl - self.x0) + self._ky0 * (yl - self.y0)
        return xb, yb - self.focus

    def beam2lab(self, xb, yb):
        yb = yb + self.focus
        xl = self._ky0 * xb + self._kx0 * yb
        yl = -self._kx0 * xb + self._ky0 * yb
        return xl + self.x0, yl + self.y0

    def rayleigh_len(self, n=1.):
        return np.pi * self.waist ** 2 * n / self.wave_len

    # ---- analytic Gaussian beam (beam.py:137-206) ---------------------------
    def field(self, *, x=0, y=0, coords=None, n=1., outfmt="field"):
        x = np.atleast_1d(x)
        y = np.atleast_1d(y)
        if coords is not None:
            if isinstance(coords, dict):
                x, y = np.array(coords["x"]), np.array(coords["y"])
            elif isinstance(coords, (list, tuple)):
                x, y = np.array(coords[0]), np.array(coords[1])
            else:
                c = np.asarray(coords)
                if c.shape[1] > c.shape[0]:
                    c = c.T
                x, y = c[:, 0], c[:, 1]
        if len(x) != len(y):
            x, y = np.meshgrid(x, y)
        xx, yy = self.lab2beam(np.squeeze(x), np.squeeze(y))
        return self._field(xx, yy, n=n, outfmt=outfmt)

    def _field(self, xx, yy, n=1., outfmt="field"):
        xx = np.asarray(xx, float)
        yy = np.asarray(yy, float)
        w0, lw = self.waist, self.wave_len
        y_r = self.rayleigh_len(n=n)
        wa = w0 * np.sqrt(1. + yy ** 2 / y_r ** 2)                 # beam width
        r = np.full_like(yy, np.inf)
        nz = yy != 0
        r[nz] = yy[nz] * (1. + y_r ** 2 / yy[nz] ** 2)             # curvature radius
        psi = np.arctan(yy / y_r) / 2                              # Gouy phase (2-D beam)
        amp = np.sqrt(w0 / wa) * np.exp(-xx ** 2 / wa ** 2)
        phase = 2 * np.pi * yy * n / lw + np.pi * xx ** 2 / (r * lw) - psi * Beam._gouy
        fld = amp * np.exp(-1j * phase)
        if outfmt == "field":
            return fld
        if outfmt == "tuple":
            return fld, amp, phase, r, wa, psi
        if outfmt == "dict":
            return {"field": fld, "amp": amp, "phase": phase, "r": r, "wa": wa, "gouy": psi}
        return None

    # ---- wavefront (beam.py:209-276) ----------------------------------------
    def _phase_and_dy(self, xx, yy, n):
        """Analytic phase(x, y) in beam coordinates and its y-derivative."""
        lw = self.wave_len
        y_r = self.rayleigh_len(n=n)
        q = yy ** 2 + y_r ** 2
        ph = 2 * np.pi * yy * n / lw + np.pi * xx ** 2 * yy / (q * lw) - np.arctan(yy / y_r) / 2 * Beam._gouy
        dph = 2 * np.pi * n / lw + np.pi * xx ** 2 * (y_r ** 2 - yy ** 2) / (q ** 2 * lw) - y_r / q / 2 * Beam._gouy
        return ph, dph

    @staticmethod
    def _iso_line(a):
        """Marching-squares vertices of the zero level of ``a`` as (row, col), ordered by column:
        one vertex on every cell edge whose end values straddle zero, placed by linear
        interpolation (the published algorithm behind skimage.measure.find_contours, beam.py:238;
        for a front that is a graph over the columns, chaining order = column order)."""
        r, c = np.nonzero(a[:-1, :] * a[1:, :] < 0)              # vertical edges
        t = a[r, c] / (a[r, c] - a[r + 1, c])
        rows, cols = [r + t], [c.astype(float)]
        r, c = np.nonzero(a[:, :-1] * a[:, 1:] < 0)              # horizontal edges
        t = a[r, c] / (a[r, c] - a[r, c + 1])
        rows.append(r.astype(float)); cols.append(c + t)
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        order = np.argsort(cols, kind="stable")
        return rows[order], cols[order]

    def _iso_line_banded(self, x, y, n, ph0, y_guess):
        """The vertices ``_iso_line`` finds on the full ``len(y) x len(x)`` phase map, from phase
        values in a narrow band of rows around the front only (the map is 10^6 nodes, the front
        touches ~3 000 of them).  The phase grows monotonically along y, so every column has
        exactly one straddled vertical edge, and a horizontal edge (row r, columns c..c+1) is
        straddled exactly when r lies between the two columns' crossing rows.  Node values come
        from the same elementwise formula, so the vertices are bit-identical to the full map's.
        Returns None when the front is not inside the rows for every column (caller widens)."""
        ny, nx = len(y), len(x)
        dy = (y[-1] - y[0]) / (ny - 1)
        r0 = np.floor((y_guess - y[0]) / dy).astype(np.int64)          # row just below the front, +-1
        if r0.min() < 2 or r0.max() > ny - 4:
            return None
        k = np.arange(-2, 4)                                            # 6 rows per column
        rr = r0[None, :] + k[:, None]
        cc = np.broadcast_to(np.arange(nx), rr.shape)
        a = self._field(x[cc], y[rr], n=n, outfmt="dict")["phase"] - ph0
        sgn = a[:-1] * a[1:] < 0
        if not np.all(sgn.sum(axis=0) == 1):
            return None
        j = np.argmax(sgn, axis=0)
        c = np.arange(nx)
        rv = rr[j, c]                                                   # crossing between rows rv and rv+1
        t = a[j, c] / (a[j, c] - a[j + 1, c])
        rows, cols = [rv + t], [c.astype(float)]
        # horizontal edges between columns c and c+1: rows in (min(rv), max(rv)] of the pair
        lo, hi = np.minimum(rv[:-1], rv[1:]), np.maximum(rv[:-1], rv[1:])
        span = int((hi - lo).max())
        if span > 0:
            kk = np.arange(1, span + 1)
            rh = lo[None, :] + kk[:, None]
            ok = rh <= hi[None, :]
            ch = np.broadcast_to(np.arange(nx - 1), rh.shape)
            rh_c = np.minimum(rh, ny - 1)
            al = self._field(x[ch], y[rh_c], n=n, outfmt="dict")["phase"] - ph0
            ar = self._field(x[ch + 1], y[rh_c], n=n, outfmt="dict")["phase"] - ph0
            ok &= al * ar < 0
            rows.append(rh[ok].astype(float)); cols.append(ch[ok] + al[ok] / (al[ok] - ar[ok]))
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        order = np.argsort(cols, kind="stable")
        return rows[order], cols[order]

    def _wavefront_contour(self, dist, ampbnd, n, banded=True):
        """The reference's finder, step for step (beam.py:209-247).  ``banded=False`` evaluates the
        full 1000 x 1000 phase map like the reference does; the default evaluates only the rows
        the front touches and returns the same vertices bit for bit."""
        w0, lw = self.waist, self.wave_len
        y_r = self.rayleigh_len(n=n)
        xmax = np.sqrt(-w0 ** 2 * (1 + (dist / y_r) ** 2) * np.log(ampbnd))        # beam.py:218
        x = np.linspace(-xmax, xmax, Beam._nfront)
        r = dist * (1 + (y_r / dist) ** 2)                                          # beam.py:223
        ph0 = 2 * np.pi * dist * n / lw - np.arctan(dist / y_r) / 2 * Beam._gouy   # beam.py:236
        y_guess = self._front_newton(x, dist, n, ph0) if banded else None
        nl = 1
        while True:                                                                 # beam.py:228-242
            y = np.linspace(dist - np.sqrt(r ** 2 - xmax ** 2) - nl * lw, dist + nl * lw, Beam._nfront)
            ind2x = interp1d(np.arange(len(x)), x)
            ind2y = interp1d(np.arange(len(y)), y)
            line = self._iso_line_banded(x, y, n, ph0, y_guess) if banded else None
            if line is None:
                xx, yy = np.meshgrid(x, y)
                line = self._iso_line(self._field(xx, yy, n=n, outfmt="dict")["phase"] - ph0)
            rows, cols = line
            if len(cols) and ind2x(cols[0]) <= -xmax:
                break
            nl += 1
            if nl > 64:
                raise RuntimeError("wavefront finder: the iso-phase line never reaches the left wall")
        return interp1d(ind2x(cols), ind2y(rows), kind="cubic")                     # beam.py:245

    def _front_newton(self, x, dist, n, ph0):
        """y(x) with phase(x, y) = ph0, by Newton's method from the spherical first guess."""
        y_r = self.rayleigh_len(n=n)
        rc = dist * (1 + (y_r / dist) ** 2)
        y = dist - (rc - np.sqrt(np.maximum(rc ** 2 - x ** 2, 0.0)))
        for _ in range(50):
            ph, dph = self._phase_and_dy(x, y, n)
            step = (ph - ph0) / dph
            y = y - step
            if np.max(np.abs(step)) < 1e-14 * max(1.0, abs(dist)):
                break
        return y

    def _wavefront(self, dist=0.003, ampbnd=0.001, n=1.):
        """Interpolant x -> y (beam coordinates) of the iso-phase line through (0, dist), sampled
        on ``_nfront`` columns out to the ``ampbnd`` amplitude level (beam.py:209-248)."""
        # In beam coordinates the front depends on the beam's shape only, not on where the antenna
        # stands: the 14 beams of one ft2.phases density step share it.
        key = (self.front_method, Beam._nfront, Beam._gouy, float(self.waist), float(self.wave_len), float(dist),
               float(ampbnd), float(np.asarray(n).reshape(-1)[0]))
        hit = Beam._front_cache.get(key)
        if hit is not None:
            return hit
        if len(Beam._front_cache) > 64:
            Beam._front_cache.clear()
        wf = self._wavefront_uncached(dist, ampbnd, n)
        Beam._front_cache[key] = wf
        return wf

    def _wavefront_uncached(self, dist, ampbnd, n):
        if self.front_method == "contour":
            return self._wavefront_contour(dist, ampbnd, n)
        if self.front_method != "newton":
            raise ValueError("Beam.front_method must be 'newton' or 'contour'")
        w0, lw = self.waist, self.wave_len
        y_r = self.rayleigh_len(n=n)
        xmax = np.sqrt(-w0 ** 2 * (1 + (dist / y_r) ** 2) * np.log(ampbnd))        # beam.py:218
        x = np.linspace(-xmax, xmax, Beam._nfront)
        ph0 = 2 * np.pi * dist * n / lw - np.arctan(dist / y_r) / 2 * Beam._gouy   # beam.py:236
        return interp1d(x, self._front_newton(x, dist, n, ph0), kind="cubic")

    def wavefront(self, dist=0.02, ampbnd=0.001, n=1.):
        wf = self._wavefront(dist, ampbnd, n=n)
        return self.beam2lab(wf.x, wf.y)

    def wavefront_r(self, dist=0.03, ampbnd=0.001, n_field=None):
        """Wavefront for the ray generator: the iso-phase line nearest to ``dist`` whose on-axis
        phase is a whole number of periods (beam.py:256-276)."""
        n = 1. if n_field is None else n_field.refractive_index(freq=self.freq)
        lw = self.wave_len
        y_r = self.rayleigh_len(n=n)
        nl = np.floor(n * dist / lw - Beam._gouy * np.arctan(dist / y_r) / 2 / 2 / np.pi)
        y = np.linspace((nl - 1) * lw, (nl + 1) * lw, 100)
        n2y = interp1d(n * y / lw - Beam._gouy * np.arctan(y / y_r) / 2 / 2 / np.pi, y)
        return self._wavefront(float(n2y(nl)), ampbnd, n=n)

    # ---- ray launch (beam.py:278-344) ---------------------------------------
    def launch_conditions(self, rnum, wave_front, n_edge):
        """Start points (lab frame), unit normals and amplitudes of ``rnum`` rays equidistant in
        arc length along ``wave_front`` (beam.py:293-331)."""
        seg = np.sqrt(np.diff(wave_front.x) ** 2 + np.diff(wave_front.y) ** 2)
        lax = np.cumsum(np.insert(seg, 0, 0))
        l2x = interp1d(lax, wave_front.x)
        x = l2x(np.linspace(0, lax[-1], rnum))
        y = wave_front(x)
        # Normals: the reference differentiates the front on a fine grid of M = lcm(rnum, _nfront)
        # arc-length-equidistant points with np.gradient and keeps every (M/rnum)-th value
        # (beam.py:306-326).  Only the kept nodes and their two neighbours enter the result, so just
        # those 3*rnum points are evaluated -- same values bit for bit, O(rnum) instead of O(lcm).
        m_fine = int(np.lcm(rnum, Beam._nfront))
        stride = int(m_fine / rnum)
        keep = np.arange(0, m_fine, stride)
        lo, hi = np.maximum(keep - 1, 0), np.minimum(keep + 1, m_fine - 1)
        step = lax[-1] / (m_fine - 1)                       # np.linspace: arange(M) * step + start, last = stop

        def fine(idx):
            t = idx * step + 0.0
            t[idx == m_fine - 1] = lax[-1]
            xb = l2x(t)
            return self.beam2lab(xb, wave_front(xb))

        x, y = self.beam2lab(x, y)
        (x_lo, y_lo), (x_hi, y_hi) = fine(lo), fine(hi)
        width = np.where((keep == 0) | (keep == m_fine - 1), 1.0, 2.0)   # one-sided at the ends, central inside
        kx, ky = -((y_hi - y_lo) / width), (x_hi - x_lo) / width
        nrm = np.sqrt(kx ** 2 + ky ** 2)
        kx, ky = kx / nrm, ky / nrm
        fs = self.field(x=x, y=y, n=n_edge, outfmt="dict")
        return x, y, kx, ky, fs["amp"], fs["phase"]

    def launch_conditions_device(self, rnum, wave_front, n_edge):
        """``launch_conditions`` on the GPU (``stb_beam_launch``): everything is a function of the ray
        index and the wavefront's node tables.  Returns CUDA tensors (x, y, kx, ky, amp); agrees
        with the host version to ~1e-12 (SciPy evaluates the same spline)."""
        import ctypes
        import torch
        from . import _cabi
        seg = np.sqrt(np.diff(wave_front.x) ** 2 + np.diff(wave_front.y) ** 2)
        lax = np.cumsum(np.insert(seg, 0, 0))
        # the cubic interp1d of the wavefront is a not-a-knot B-spline; build it through the public API
        # (bit-identical knots and coefficients to the interp1d object's own)
        from scipy.interpolate import make_interp_spline
        spl = make_interp_spline(np.asarray(wave_front.x, float), np.asarray(wave_front.y, float), k=3, check_finite=False)
        dev = torch.device("cuda", torch.cuda.current_device())
        tabs = [torch.from_numpy(np.ascontiguousarray(v, np.float64).ravel()).to(dev)
                for v in (lax, wave_front.x, spl.t, spl.c)]
        out = [torch.empty(int(rnum), dtype=torch.float64, device=dev) for _ in range(5)]
        a = _cabi.StbBeam()
        a.struct_size = ctypes.sizeof(_cabi.StbBeam)
        a.n, a.nfront = int(rnum), int(Beam._nfront)
        a.lax, a.wf_x, a.nw = tabs[0].data_ptr(), tabs[1].data_ptr(), int(tabs[0].shape[0])
        a.t, a.c, a.nt = tabs[2].data_ptr(), tabs[3].data_ptr(), int(tabs[2].shape[0])
        a.x0, a.y0, a.kx0, a.ky0, a.focus = float(self.x0), float(self.y0), float(self._kx0), float(self._ky0), float(self.focus)
        a.waist, a.rayleigh = float(self.waist), float(self.rayleigh_len(n=n_edge))
        a.x, a.y, a.kx, a.ky, a.amp = (t.data_ptr() for t in out)
        _cabi.check(_cabi.lib.stb_set_device(dev.index or 0))
        _cabi.check(_cabi.lib.stb_beam_launch(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        torch.cuda.current_stream(dev).synchronize()      # the tables go out of scope here
        return tuple(out)

    def rays(self, rnum=None, wave_front=None, dist=None, ampbnd=None, n_field=None, device_launch=None, **trace_kw):
        """Generate and trace the beam's rays: one kernel launch for the whole fan.  For >= 65536
        rays (or ``device_launch=True``) the launch conditions are generated on the GPU as well."""
        rnum = Beam._nray if rnum is None else rnum
        dist = Beam._dist if dist is None else dist
        ampbnd = Beam._ampbnd if ampbnd is None else ampbnd
        if n_field is None:
            raise ValueError("rays() needs a density field (the reference dereferences it too, beam.py:328)")
        if wave_front is None:
            wave_front = self.wavefront_r(dist=dist, ampbnd=ampbnd, n_field=n_field)
        n_edge = n_field.refractive_index(freq=self.freq)
        if device_launch is None:
            device_launch = rnum >= 65536
        if device_launch:
            x, y, kx, ky, amps = (t.cpu().numpy() for t in self.launch_conditions_device(rnum, wave_front, n_edge))
            phase = None                                # returned "for check" only by the reference (beam.py:331)
        else:
            x, y, kx, ky, amps, phase = self.launch_conditions(rnum, wave_front, n_edge)
        k = 100 if self.length_unit == "m" else 1     # tracing works in CGS (beam.py:335-338)
        self.bundle = trace_rays(n_field, x * k, y * k, kx, ky, self.freq, amp=amps, want_extras=True, **trace_kw)
        self._ray_cache = None
        return x, y, kx, ky, amps, phase, dist

    @property
    def _rays(self):
        """Reference-shaped Ray objects (beam.py:340-342), materialised from HBM on first use."""
        if self._ray_cache is None:
            n = 0 if self.bundle is None else len(self.bundle)
            self._ray_cache = np.array([self.bundle.ray(r) for r in range(n)], dtype=object)
        return self._ray_cache

    @_rays.setter
    def _rays(self, value):
        self._ray_cache = value

    # ---- field on the receiver line from the rays (beam.py:347-444) ----------
    def field_r(self, y, axis="y"):
        if self.length_unit == "m":
            y = y * 100
        rays = self._rays
        ph, xout, amp0 = [], [], []
        pnt = []
        keep = np.ones(len(rays), dtype=bool)
        for j, r in enumerate(rays):
            ind = r.addpoint(y, axis=axis)
            if not hasattr(ind, "__len__") or len(ind) != 1:
                keep[j] = False                        # no crossing, or a ray that crosses twice
                continue
            amp0.append(r.amp)
            pnt.append(int(ind[0]))
            ph.append(r.ph[ind[0]])
            xout.append(r.x[ind[0]])
        rays = rays[keep]
        self._rays = rays
        ph, xout, amp0, pnt = np.array(ph), np.array(xout), np.array(amp0, float), np.array(pnt, int)
        amp1 = amp0.copy()
        n = len(rays)

        # sort the rays into "pages" of monotonically increasing receiver x (caustics fold the beam)
        pages = _pages(xout)

        field = np.zeros(n, dtype=complex)
        for pg in pages:
            d0 = np.zeros((len(pg), 2))
            d1 = np.zeros((len(pg), 2))
            for a, m in enumerate(pg):
                rm, pi = rays[m], pnt[m]
                for b, o in enumerate((pg[a - 1], pg[(a + 1) % len(pg)])):
                    ro = rays[o]
                    d0[a, b] = np.hypot(rm.x[0] - ro.x[0], rm.y[0] - ro.y[0])
                    # the neighbour's point at the same eikonal phase
                    xo = interp1d(ro.ph, ro.x, fill_value="extrapolate")(rm.ph[pi])
                    yo = interp1d(ro.ph, ro.y, fill_value="extrapolate")(rm.ph[pi])
                    d1[a, b] = np.hypot(xo - rm.x[pi], yo - rm.y[pi])
            d0, d1 = d0.reshape(-1), d1.reshape(-1)
            d0[0], d0[-1] = d0[1], d0[-2]
            d1[0], d1[-1] = d1[1], d1[-2]
            # ray-tube divergence: in 2-D the amplitude falls as 1/sqrt(separation)
            with np.errstate(divide="ignore", invalid="ignore"):
                a1 = np.sqrt(d0.reshape(-1, 2).sum(1) / d1.reshape(-1, 2).sum(1))
            a1[np.isnan(a1)] = amp0[pg][np.isnan(a1)]
            amp1[pg] = a1
            field[pg] += amp0[pg] * a1 * np.exp(-1j * ph[pg])

        order = np.argsort(xout)
        xout, field, amp0, amp1, ph, pnt = xout[order], field[order], amp0[order], amp1[order], ph[order], pnt[order]
        hole = np.where(field == 0)[0]
        if len(hole):
            amp1[hole] = (np.roll(amp1, -1) + np.roll(amp1, 1))[hole] / 2
            amp1[np.isnan(amp1)] = amp0[np.isnan(amp1)]
            field[hole] = amp0[hole] * amp1[hole] * np.exp(-1j * ph[hole])
        return field, xout, amp0, amp1, ph, pnt


def _pages(xout):
    """Split ray indices into "pages" of increasing receiver coordinate (beam.py:374-398).

    The reference walks the rays once per page: a page starts at a seed; a later ray joins it when
    it lies beyond the page's current last member and belongs to no page yet; the first ray that
    does NOT lie beyond it seeds the next page.  Pages with fewer than two rays are dropped.
    The nearest-cell scheme makes the launch -> receiver map jump at cell edges, so dense beams
    have hundreds of tiny pages and the walk is O(rays x pages): it runs natively
    (``stb_field_pages``, host code of the C library)."""
    import ctypes
    from . import _cabi
    xout = np.ascontiguousarray(xout, np.float64)
    n = len(xout)
    if n >= 2 and bool(np.all(xout[1:] > xout[:-1])):
        return [np.arange(n)]                  # no caustic, no jumps: the whole beam is one page
    offsets = np.zeros(n + 2, np.int64)
    members = np.zeros(2 * n + 2, np.int64)
    npages = ctypes.c_int64()
    _cabi.check(_cabi.lib.stb_field_pages(xout.ctypes.data, n, offsets.ctypes.data, members.ctypes.data, ctypes.byref(npages)))
    return [members[offsets[k]:offsets[k + 1]].copy() for k in range(npages.value)]


def _pages_loop(xout):
    """Literal restatement of beam.py:374-398 (kept as the checker of ``_pages`` in the tests)."""
    n = len(xout)
    pages = [[0]] if n else []
    cur = 0
    used = {0}
    while cur < len(pages):
        for j in range(pages[cur][-1] + 1, n):
            if xout[j] > xout[pages[cur][-1]]:
                if j not in used:
                    pages[cur].append(j)
                    used.add(j)
            elif cur + 1 == len(pages):
                pages.append([j])
                used.add(j)
        cur += 1
    return [np.array(pg) for pg in pages if len(pg) >= 2]


def field_r_device(beam: Beam, y, axis="y"):
    """``Beam.field_r`` (beam.py:347-444) without materialising a single Ray: the crossing of the
    receiver line comes from the fused ``cross`` output of the trace, the neighbour rays' points at
    equal eikonal phase from one more launch (``lookup_ph``); only O(rays) numbers reach the host.
    Returns the same tuple as ``field_r``."""
    from . import sidtrace as _st
    from .ray import _device_field
    if beam.length_unit == "m":
        y = y * 100
    bun = beam.bundle
    nf = bun.n_field
    fld = _device_field(nf)
    kw = dict(L=int(bun.L), st=bun.st, dx=float(nf.dx[0]), dy=float(nf.dy[0]))
    res = bun.result
    if res is None or res.cross_count is None or bun.receiver != (float(y), axis):
        res = _st.trace_batch(fld, bun.xg0, bun.yg0, bun.kx0, bun.ky0, cross=(float(y), axis),
                              x0_phys=bun.x0, y0_phys=bun.y0, **kw)
    keep = np.flatnonzero(np.asarray(res.cross_count) == 1)     # exactly one crossing (beam.py:361-364)
    cs = np.asarray(res.cross_state)[:, keep]
    xc, yc, ph = cs[0], cs[1], cs[5]
    xout = xc.copy()                                             # the reference reads r.x[ind] for either axis
    amp0 = np.asarray(bun.amp, float)[keep]
    amp1 = amp0.copy()
    pnt = np.asarray(res.cross_index)[keep]
    x0, y0 = bun.x0[keep], bun.y0[keep]
    n = len(keep)
    pages = _pages(xout)

    # every ray of a page is asked for its point at the phase of its right-hand (A) and
    # left-hand (B) page neighbour's crossing
    targets = np.full((2, n), np.nan)
    for pg in pages:
        targets[0, pg] = ph[np.roll(pg, -1)]
        targets[1, pg] = ph[np.roll(pg, 1)]
    sub = [np.ascontiguousarray(a[keep]) for a in (bun.xg0, bun.yg0, bun.kx0, bun.ky0)]
    look = _st.trace_batch(fld, *sub, cross=(float(y), axis), x0_phys=np.ascontiguousarray(x0),
                           y0_phys=np.ascontiguousarray(y0), lookup_ph=np.ascontiguousarray(targets), **kw)
    lx = np.asarray(look.lookup_xy)

    field = np.zeros(n, dtype=complex)
    for pg in pages:
        left, right = np.roll(pg, 1), np.roll(pg, -1)
        d0 = np.stack([np.hypot(x0[pg] - x0[left], y0[pg] - y0[left]),
                       np.hypot(x0[pg] - x0[right], y0[pg] - y0[right])], 1).reshape(-1)
        # left neighbour sampled at MY phase = its A target; right neighbour = its B target
        d1 = np.stack([np.hypot(lx[0, left] - xc[pg], lx[1, left] - yc[pg]),
                       np.hypot(lx[2, right] - xc[pg], lx[3, right] - yc[pg])], 1).reshape(-1)
        d0[0], d0[-1] = d0[1], d0[-2]
        d1[0], d1[-1] = d1[1], d1[-2]
        with np.errstate(divide="ignore", invalid="ignore"):
            a1 = np.sqrt(d0.reshape(-1, 2).sum(1) / d1.reshape(-1, 2).sum(1))
        a1[np.isnan(a1)] = amp0[pg][np.isnan(a1)]
        amp1[pg] = a1
        field[pg] += amp0[pg] * a1 * np.exp(-1j * ph[pg])

    order = np.argsort(xout)
    xout, field, amp0, amp1, ph, pnt = xout[order], field[order], amp0[order], amp1[order], ph[order], pnt[order]
    hole = np.where(field == 0)[0]
    if len(hole):
        amp1[hole] = (np.roll(amp1, -1) + np.roll(amp1, 1))[hole] / 2
        amp1[np.isnan(amp1)] = amp0[np.isnan(amp1)]
        field[hole] = amp0[hole] * amp1[hole] * np.exp(-1j * ph[hole])
    return field, xout, amp0, amp1, ph, pnt


def receiver_field_device(beam: Beam, y, axis="y"):
    """``Beam.field_r`` (beam.py:347-444) with every O(rays) array on the GPU: the crossing of the receiver
    line comes from the fused ``cross`` output of the trace, the page structure from
    ``stb_field_pages_links`` (first-fit by bisection; skipped when the receiver coordinate is monotonic
    in the launch order -- one page), the neighbour rays' points at equal eikonal phase from the lookup
    launch, ray-tube amplitudes and the complex field from ``stb_tube_field``, the rays of dropped pages
    from ``stb_hole_fill``.  Returns a dict of CUDA tensors sorted by receiver coordinate:
    field [n, 2] (re, im), xout, amp0, amp1, ph, pnt."""
    import ctypes
    import torch
    from . import _cabi, sidtrace as _st
    from .ray import _device_field
    if beam.length_unit == "m":
        y = y * 100
    bun = beam.bundle
    nf = bun.n_field
    fld = _device_field(nf)
    dev = torch.device("cuda", torch.cuda.current_device())
    # the cell table whose fast axis follows the beam (same cells, same bits; STB_F_YMAJOR): the FT-2 antennas shoot
    # along y, so their rays find the next cells in the same cache line of the y-major copy
    cells = "y" if abs(beam.ky0) > abs(beam.kx0) else "x"
    kw = dict(L=int(bun.L), st=bun.st, dx=float(nf.dx[0]), dy=float(nf.dy[0]), cells=cells)

    def up(a):
        return a if hasattr(a, "is_cuda") else torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    if bun.dev is not None:
        launch, x0p, y0p, amp = bun.dev["launch"], bun.dev["x0"], bun.dev["y0"], bun.dev["amp"]
    else:
        launch = [up(a) for a in (bun.xg0, bun.yg0, bun.kx0, bun.ky0)]
        x0p, y0p, amp = up(bun.x0), up(bun.y0), up(np.asarray(bun.amp, float))
    res = bun.result
    if res is None or res.cross_count is None or bun.receiver != (float(y), axis) or not hasattr(res.cross_count, "is_cuda"):
        res = _st.trace_batch(fld, *launch, cross=(float(y), axis), x0_phys=x0p, y0_phys=y0p, **kw)
    keep = torch.nonzero(res.cross_count == 1).squeeze(1)            # exactly one crossing (beam.py:361-364)
    n = int(keep.numel())
    cs = res.cross_state.index_select(1, keep)
    xc, yc, ph = cs[0].contiguous(), cs[1].contiguous(), cs[5].contiguous()
    amp0 = amp.index_select(0, keep)
    pnt = res.cross_index.index_select(0, keep)
    x0k, y0k = x0p.index_select(0, keep), y0p.index_select(0, keep)
    i32 = dict(dtype=torch.int32, device=dev)
    if n < 2 or bool((xc[1:] > xc[:-1]).all()):                     # no caustic, no jumps: one page
        pred = torch.arange(-1, n - 1, **i32)
        succ = torch.arange(1, n + 1, **i32)
        if n:
            succ[-1] = -1
        page_size = torch.full((n,), n, **i32)
    else:
        hx = xc.cpu().numpy()
        hp, hs, hz = (np.empty(n, np.int32) for _ in range(3))
        npages = ctypes.c_int64()
        _cabi.check(_cabi.lib.stb_field_pages_links(hx.ctypes.data, n, hp.ctypes.data, hs.ctypes.data, hz.ctypes.data,
                                                    ctypes.byref(npages)))
        pred, succ, page_size = (torch.from_numpy(a).to(dev) for a in (hp, hs, hz))
    # every ray is asked for its point at the phase of its right-hand (A) and left-hand (B) page
    # neighbour's crossing (a page end asks for its own phase; that answer is never used)
    own = torch.arange(n, device=dev)
    tA = ph[torch.where(succ >= 0, succ.long(), own)]
    tB = ph[torch.where(pred >= 0, pred.long(), own)]
    amp1 = torch.empty(n, dtype=torch.float64, device=dev)
    field = torch.empty((n, 2), dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    if n:
        look = _st.trace_batch(fld, *[a.index_select(0, keep) for a in launch], cross=(float(y), axis), x0_phys=x0k, y0_phys=y0k,
                               lookup_ph=torch.stack([tA, tB]).contiguous(), **kw)
        a = _cabi.StbTube()
        a.struct_size = ctypes.sizeof(_cabi.StbTube)
        a.n = n
        a.pred, a.succ, a.page_size = pred.data_ptr(), succ.data_ptr(), page_size.data_ptr()
        a.x0, a.y0, a.xc, a.yc, a.ph = x0k.data_ptr(), y0k.data_ptr(), xc.data_ptr(), yc.data_ptr(), ph.data_ptr()
        a.amp0, a.lookup_xy = amp0.data_ptr(), look.lookup_xy.data_ptr()
        a.amp1, a.field = amp1.data_ptr(), field.data_ptr()
        _cabi.check(_cabi.lib.stb_tube_field(ctypes.byref(a), stream))
    order = torch.argsort(xc)
    xout, field, amp0, amp1, ph, pnt = (t.index_select(0, order) for t in (xc, field, amp0, amp1, ph, pnt))
    field = field.contiguous()
    amp1f = torch.empty_like(amp1)
    _cabi.check(_cabi.lib.stb_hole_fill(n, amp0.data_ptr(), ph.data_ptr(), amp1.data_ptr(), amp1f.data_ptr(), field.data_ptr(), stream))
    return {"field": field, "xout": xout, "amp0": amp0, "amp1": amp1f, "ph": ph, "pnt": pnt}


def _gauss_fit_device(x, y, p0, stream):
    """Least-squares fit of amp * exp(-(x - xc)^2 / (2 sigma^2)) to (x, y) -- what the reference asks of
    scipy.optimize.curve_fit (beam.py:463-465) -- with the data on the GPU: MINPACK's lmdif restated over
    reductions (``lmfit.lmdif3``; one ``stb_gauss_fit_sums`` launch per trial point, forward-difference
    Jacobian with MINPACK's steps, the 3 x 3 algebra on the host), so that it stops where curve_fit stops
    even on the flat cost surfaces of folded beams.  Beams of fewer than 8192 rays, and fits whose Jacobian
    is nearly rank deficient (few rays inside the Gaussian: the normal equations lose what MINPACK's pivoted
    QR keeps), go through SciPy's curve_fit on a host copy -- the reference's own call.  None when the fit
    fails (curve_fit raises)."""
    import ctypes
    import torch
    from . import _cabi
    from .lmfit import ILL_CONDITIONED, lmdif3
    n = int(x.shape[0])
    out = torch.empty(10, dtype=torch.float64, device=x.device)

    def sums(p, h):
        hs = (ctypes.c_double * 3)(*[float(v) for v in h])
        _cabi.check(_cabi.lib.stb_gauss_fit_sums(x.data_ptr(), y.data_ptr(), n, float(p[0]), float(p[1]), float(p[2]),
                                                 hs, out.data_ptr(), stream))
        s = out.cpu().numpy()
        A = np.array([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]])
        return A, s[6:9].copy(), float(s[9])

    if n > _FIT_ON_HOST_BELOW:
        p, info = lmdif3(sums, p0)
        if info != ILL_CONDITIONED:
            return p if info in (1, 2, 3, 4) and np.all(np.isfinite(p)) else None
    # a small beam, or one with too few rays inside the fitted Gaussian for the normal equations: the
    # reference's own routine on a host copy of (x, |field|) -- O(n) bytes once, no device work involved
    try:
        p, _ = curve_fit(lambda xa, amp, xc, sigma: amp * np.exp(-(xa - xc) ** 2 / 2 / sigma ** 2),
                         x.cpu().numpy(), y.cpu().numpy(), p0)
        return p
    except Exception:
        return None


_FIT_ON_HOST_BELOW = 8192     # rays: below this the D2H copy + MINPACK cost less than the reduction launches


def _cubic_at(x, ph, xt, halfwidth=128):
    """``interp1d(x, ph, kind='cubic')(xt)`` (beam.py:453) from a window of nodes around ``xt``: the
    not-a-knot spline's dependence on a node decays by a factor ~(2 - sqrt 3) per node in between, so 128
    nodes either side leave nothing in double precision -- and an O(n) banded solve for ONE value is
    avoided.  ``x``, ``ph``: CUDA tensors or arrays, x increasing.  Raises ValueError outside the range."""
    from scipy.interpolate import make_interp_spline
    n = int(x.shape[0])
    if not (float(x[0]) <= xt <= float(x[-1])):
        raise ValueError("A value in x_new is outside the interpolation range.")
    if hasattr(x, "is_cuda"):
        import torch
        i = int(torch.searchsorted(x, torch.tensor([xt], dtype=x.dtype, device=x.device))[0])
    else:
        i = int(np.searchsorted(x, xt))
    lo, hi = max(0, i - halfwidth), min(n, i + halfwidth)
    if hi - lo < 8:
        lo, hi = 0, n
    xs, ps = x[lo:hi], ph[lo:hi]
    if hasattr(xs, "is_cuda"):
        xs, ps = xs.cpu().numpy(), ps.cpu().numpy()
    return make_interp_spline(np.asarray(xs, float), np.asarray(ps, float), k=3, check_finite=False)(xt)


def field_result_device(beam_tx: Beam, beam_rx: Beam, want_arrays: bool = True):
    """``field_result`` (beam.py:447-472) on the GPU: receiver-line field of the traced TX beam
    (``receiver_field_device``) times the analytic RX pattern, |f|-weighted centre and the complex signal
    in one launch (``stb_overlap``), the Gaussian fit by MINPACK's lmdif over device reductions, the
    cubic phase at the fitted centre from a window of nodes.  Only O(1) numbers cross PCIe unless
    ``want_arrays`` (then field and x come back as the reference returns them)."""
    import ctypes
    import torch
    from . import _cabi
    r = receiver_field_device(beam_tx, beam_rx.y0)
    x, ph, ftx = r["xout"].contiguous(), r["ph"].contiguous(), r["field"]
    n = int(x.shape[0])
    dev = x.device
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    field = torch.empty((n, 2), dtype=torch.float64, device=dev)
    absf = torch.empty(n, dtype=torch.float64, device=dev)
    sums = torch.empty(8, dtype=torch.float64, device=dev)
    a = _cabi.StbOverlap()
    a.struct_size = ctypes.sizeof(_cabi.StbOverlap)
    a.n = n
    a.x, a.field_tx = x.data_ptr(), ftx.data_ptr()
    k = 100 if beam_rx.length_unit == "m" else 1
    a.y_rx = float(beam_rx.y0)
    a.bx0, a.by0, a.bkx, a.bky, a.focus = float(beam_rx.x0), float(beam_rx.y0), float(beam_rx._kx0), float(beam_rx._ky0), float(beam_rx.focus)
    a.waist, a.wave_len, a.rayleigh, a.gouy = float(beam_rx.waist), float(beam_rx.wave_len), float(beam_rx.rayleigh_len(n=1.)), float(Beam._gouy)
    a.field, a.absf, a.sums = field.data_ptr(), absf.data_ptr(), sums.data_ptr()
    del k
    _cabi.check(_cabi.lib.stb_overlap(ctypes.byref(a), stream))
    s = sums.cpu().numpy()
    x_max = s[1] / s[0]
    signal = complex(s[2], s[3])
    fmax = float(np.frombuffer(s[4:5].tobytes(), np.float64)[0])
    props = _gauss_fit_device(x, absf, (fmax, x_max, beam_tx.waist), stream) if n >= 3 else None
    try:
        if props is None:
            raise RuntimeError("the Gaussian fit did not converge")
        ph_max = _cubic_at(x, ph, float(props[1]))
    except Exception:
        ph_max = _cubic_at(x, ph, float(x_max))
        props = None
    if want_arrays:
        f = field.cpu().numpy()
        return f[:, 0] + 1j * f[:, 1], x.cpu().numpy(), x_max, ph_max, signal, props
    return None, None, x_max, ph_max, signal, props


def field_result(beam_tx: Beam, beam_rx: Beam, device: bool = False):
    """Overlap of the traced TX field with the analytic RX pattern on the receiver line:
    complex interferometer signal and phase at the maximum (beam.py:447-472).  ``device=True``
    takes the receiver-line field from ``field_r_device`` (no Ray objects); ``device="gpu"`` runs the whole
    of it on the GPU (``field_result_device``)."""
    if device == "gpu":
        return field_result_device(beam_tx, beam_rx)
    field_tx, x, _, _, ph, pnt = field_r_device(beam_tx, beam_rx.y0) if device else beam_tx.field_r(beam_rx.y0)
    field = field_tx * beam_rx.field(x=x, y=beam_rx.y0)
    x_max = np.sum(np.abs(field) * x) / np.sum(np.abs(field))
    x2ph = interp1d(x, ph, kind="cubic")
    dist = ((np.roll(x, -1) - x) + (x - np.roll(x, 1))) / 2
    dist[0], dist[-1] = dist[1], dist[-2]
    signal = np.sum(field * dist)
    try:
        props, _ = curve_fit(lambda xa, amp, xc, sigma: amp * np.exp(-(xa - xc) ** 2 / 2 / sigma ** 2), x, np.abs(field),
                             (np.abs(field).max(), x_max, beam_tx.waist))
        ph_max = x2ph(props[1])
    except Exception:
        ph_max = x2ph(x_max)
        props = None
    return field, x, x_max, ph_max, signal, props
