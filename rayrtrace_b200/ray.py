"""Caller adapter of the hot path: the reference's ``Ray`` (ray.py:15-126) and
the batched launch that replaces the serial ``for ...: Ray(...)`` loop of
beam.py:340-342.

``Ray(...)`` keeps the reference constructor, class-level settings
(``MaxStep``, ``kStep``, ``traceType``) and attributes (``x, y, kx, ky, st,
k, lax, ph``) and traces through ``sidtrace.trace`` on the GPU.

``trace_rays(...)`` takes arrays of launch conditions, applies the same launch
processing (ray.py:28-68) vectorised, and traces the whole set in ONE kernel
launch; trajectories stay in HBM and are read back per ray on demand.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
from scipy import constants as const

from . import sidtrace as _st

from .launch import _DEFAULT_FREQ, _vacuum_k, launch_states  # noqa: F401  (pure NumPy part, re-exported)


class RayOutsideGrid(IndexError, ValueError):
    """A ray launched outside the density grid takes no step; the reference then fails with an
    IndexError at ray.py:105 (``self.x[0]`` of an empty array).  Subclasses ValueError as well, which
    is what earlier versions of this module raised."""


class Ray:
    """One ray, traced at construction when ``n_field`` is given (ray.py:15-111)."""

    MaxStep = 100000   # ray.py:11  array length handed to trace (cap = MaxStep-1 steps)
    kStep = 1          # ray.py:12  multiplier of n_field.st
    traceType = "c"    # ray.py:13  only the native path exists here (it runs on the GPU)

    def __init__(self, x0=0, y0=0, kx0=0, ky0=1, freq=_DEFAULT_FREQ, amp=1, n_field=None):
        self.len_unit = "cm"
        L = int(Ray.MaxStep)
        x = np.zeros(L)
        y = np.zeros(L)
        kx = np.zeros(L)
        ky = np.zeros(L)
        x[0], y[0] = x0, y0
        kx0, ky0, self.freq = _vacuum_k(kx0, ky0, freq)
        kx[0], ky[0] = kx0, ky0
        self.amp = amp
        self.n_field = n_field
        self.x, self.y, self.kx, self.ky = x, y, kx, ky
        if n_field is None:
            return
        if Ray.traceType != "c":
            raise NotImplementedError(
                "only traceType 'c' exists: the pure-Python twin (ray.py:85-96) is a CPU path; "
                "see oracle/ for the checker")

        # local refractive index scales the vacuum wave vector (ray.py:62-64)
        refr = n_field.refractive_index(freq=self.freq, x=x[0], y=y[0])
        refr = float(np.asarray(refr).reshape(-1)[0])
        kx[0] = kx[0] * refr
        ky[0] = ky[0] * refr
        # physical -> grid coordinates (ray.py:67-68)
        x[0] = n_field.x2grid(x[0])
        y[0] = n_field.y2grid(y[0])
        dx = n_field.dx[0]   # equidistant grid assumed (ray.py:70-73)
        dy = n_field.dy[0]
        st = n_field.st * Ray.kStep

        i = _st.trace(x, y, kx, ky, st, dx, dy, n_field.nx, n_field.ny,
                      np.ascontiguousarray(n_field.ax), np.ascontiguousarray(n_field.ay))
        if i == 0:
            raise RayOutsideGrid(f"ray starts outside the density grid: ({x0}, {y0})")  # reference: IndexError at ray.py:105

        # keep [:i] -- the first out-of-domain state is dropped (ray.py:99-102)
        x, y, kx, ky = x[:i], y[:i], kx[:i], ky[:i]
        # grid -> physical (ray.py:105-106)
        self.x = x0 + (x - x[0]) * n_field.dx[0]
        self.y = y0 + (y - y[0]) * n_field.dy[0]
        self.kx, self.ky = kx, ky
        # per-point step; the last point starts no segment (ray.py:110-111)
        self.st = np.full_like(self.x, st)
        self.st[-1] = 0

    # ray.py:113-126
    @property
    def k(self):
        return np.sqrt(self.kx ** 2 + self.ky ** 2)

    @property
    def lax(self):
        k = self.k
        kavr = (k[:-1] + k[1:]) / 2
        return np.cumsum(np.insert(self.st[:-1] * kavr, 0, 0))

    @property
    def ph(self):
        k = self.k
        kavr = (k[:-1] + k[1:]) / 2
        dl = np.diff(self.lax)
        return np.cumsum(np.insert(dl * kavr, 0, 0))


    # ray.py:131-224
    def addpoint(self, val, axis="y"):
        """Insert the exact point(s) where the ray crosses the line ``y = val`` (or ``x = val``)
        and return their indices, with the reference's semantics (ray.py:131-224):

        * ``val`` already among the coordinates -> indices of those points, nothing inserted;
        * sign changes of ``coord - val`` -> one inserted point per crossing segment, found by
          solving ``k*s + a*s**2/2 = val - coord`` with the constant acceleration of the
          segment's start cell (smallest non-negative root, small-``a`` fallback);
        * no crossing but the last point is heading for the line -> extrapolated point appended;
        * otherwise ``nan``.
        The arrays ``x, y, kx, ky, st`` are updated in place of the old ones.  (The fused device
        version, ``trace_batch(cross=...)``, reports the first crossing without materialising rays.)"""
        c = self.y if axis == "y" else self.x
        kc_end = self.ky[-1] if axis == "y" else self.kx[-1]
        d = c - val
        if np.any(d == 0):
            return np.flatnonzero(d == 0)
        if np.any(d < 0) and np.any(d > 0):
            seg = np.flatnonzero(d[:-1] * d[1:] < 0)
        elif (val - c[-1]) * kc_end > 0:
            seg = np.array([len(c) - 1])
        else:
            return np.nan

        nf = self.n_field
        xi = np.trunc(nf.x2grid(self.x[seg])).astype(int)
        yi = np.trunc(nf.y2grid(self.y[seg])).astype(int)
        acc_x, acc_y = nf.ax[yi, xi], nf.ay[yi, xi]
        start = [np.array(a[seg]) for a in (self.x, self.y, self.kx, self.ky, self.st)]
        inserted = []
        shift = 0
        for xx, yy, kx, ky, st1, ax, ay, ind in zip(*start, acc_x, acc_y, seg):
            c0, k, a = (yy, ky, ay) if axis == "y" else (xx, kx, ax)
            s = st1
            if a != 0:
                sq = np.sqrt(k ** 2 + 2 * a * (val - c0))
                roots = np.array(((-k - sq) / a, (-k + sq) / a))
                s = roots[roots >= 0].min()
                if s == 0 and val != c0 and k != 0:     # tiny |a|: the non-zero root can be lost
                    s = (val - c0) / k
            elif k != 0:
                s = (val - c0) / k
            if s == 0:
                continue
            at = ind + shift + 1
            inserted.append(at)
            if s != st1:
                self.st[at - 1] = s
                self.st = np.insert(self.st, at, st1 - s if at != len(self.x) else 0)
                self.kx = np.insert(self.kx, at, kx + ax * s)
                self.ky = np.insert(self.ky, at, ky + ay * s)
                self.x = np.insert(self.x, at, xx + kx * s + ax / 2 * s ** 2)
                self.y = np.insert(self.y, at, yy + ky * s + ay / 2 * s ** 2)
                shift += 1
        return np.array(inserted, dtype=int)


# --------------------------------------------------------------------------
@dataclass
class RayBundle:
    """Result of ``trace_rays``: per-ray summaries on the host, trajectories in HBM."""
    n_field: object
    x0: np.ndarray            # physical launch positions
    y0: np.ndarray
    freq: float
    st: float
    amp: np.ndarray
    xg0: np.ndarray           # launch state handed to the kernel (grid units / cm^-1)
    yg0: np.ndarray
    kx0: np.ndarray
    ky0: np.ndarray
    nsteps: np.ndarray        # int32, = len(Ray.x) of the reference
    end_state: np.ndarray     # [4, n] first out-of-domain state (grid units), dropped by Ray
    result: object = None     # sidtrace.BatchResult (traj is a CUDA tensor or None)
    L: int = 0                # states per ray the launch was sized for
    receiver: object = None   # (val, axis) of the fused crossing output, if any
    dev: object = None        # CUDA copies kept for the device-side consumers: launch [4 tensors], x0, y0, amp

    def __len__(self):
        return len(self.nsteps)

    def ray(self, r: int) -> "Ray":
        """Materialise ray r as a reference-shaped Ray object (reads its
        trajectory back from HBM)."""
        if self.result is None or self.result.traj is None:
            raise RuntimeError("trace_rays(..., keep_traj=False) kept no trajectories")
        n = int(self.nsteps[r])
        t = self.result.ray_states(r, n)
        t = t.cpu().numpy() if hasattr(t, "cpu") else np.ascontiguousarray(t)
        out = Ray.__new__(Ray)
        out.len_unit = "cm"
        out.freq = self.freq
        out.amp = self.amp[r]
        out.n_field = self.n_field
        xg, yg = t[:, 0], t[:, 1]
        out.x = self.x0[r] + (xg - xg[0]) * self.n_field.dx[0] if n else xg.copy()
        out.y = self.y0[r] + (yg - yg[0]) * self.n_field.dy[0] if n else yg.copy()
        out.kx = np.ascontiguousarray(t[:, 2])
        out.ky = np.ascontiguousarray(t[:, 3])
        out.st = np.full_like(out.x, self.st)
        if n:
            out.st[-1] = 0
        return out


def _device_field(n_field):
    """HBM-resident field of an Nfield-like object: this repo's Nfield has device_field();
    the reference's own den.Nfield (duck-typed: ax, ay, nx, ny) gets one attached."""
    if hasattr(n_field, "device_field"):
        return n_field.device_field()
    f = getattr(n_field, "_stb_device_field", None)
    if f is None:
        f = _st.Field(np.ascontiguousarray(n_field.ax, np.float64), np.ascontiguousarray(n_field.ay, np.float64),
                      n_field.nx, n_field.ny)
        n_field._stb_device_field = f
    return f


def launch_states_device(n_field, x0, y0, kx0, ky0, freq):
    """``launch_states`` on the GPU (one kernel, ``stb_launch_states``): the start points are mapped
    to grid units bit-identically to the host path; the refractive index comes from the same SciPy
    spline coefficients evaluated on the device (agrees to rounding).  Returns CUDA tensors
    (xg0, yg0, kx, ky) and freq.  For beams of 10^5-10^8 rays, where SciPy's point-wise spline
    evaluation (0.7 s per 10^6 rays) would dwarf the trace itself."""
    import ctypes
    import torch
    from . import _cabi
    from .nfield import critical_den
    if freq == _DEFAULT_FREQ:
        # same rule as launch_states (ray.py:34-51): the default frequency is a sentinel only for |k| > 2
        def _kmax(a, b):
            if type(a).__module__.startswith("torch") or type(b).__module__.startswith("torch"):
                import torch as _t
                a, b = (_t.as_tensor(v, dtype=_t.float64) for v in (a, b))
                return float(_t.sqrt(a.cpu() ** 2 + b.cpu() ** 2).max())
            return float(np.sqrt(np.asarray(a, np.float64) ** 2 + np.asarray(b, np.float64) ** 2).max())
        if _kmax(kx0, ky0) > 2:
            raise ValueError("a batch must share one frequency: pass freq explicitly (ray.py:47-51 infers it per ray)")
    tabs = n_field.device_launch_tables()
    dev = tabs["x"].device

    def up(a, n=None):
        if type(a).__module__.startswith("torch"):
            return a.to(dev, torch.float64).contiguous()
        a = np.asarray(a, np.float64)
        if n is not None and a.shape != (n,):
            a = np.broadcast_to(a, (n,)).copy()
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    x0d = up(np.atleast_1d(x0) if not type(x0).__module__.startswith("torch") else x0)
    n = int(x0d.shape[0])
    y0d, kxd, kyd = up(y0, n), up(kx0, n), up(ky0, n)
    out = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
    a = _cabi.StbLaunch()
    a.struct_size = ctypes.sizeof(_cabi.StbLaunch)
    a.n = n
    a.x0_phys, a.y0_phys, a.kx_dir, a.ky_dir = x0d.data_ptr(), y0d.data_ptr(), kxd.data_ptr(), kyd.data_ptr()
    a.x_nodes, a.y_nodes, a.nx, a.ny = tabs["x"].data_ptr(), tabs["y"].data_ptr(), int(n_field.nx), int(n_field.ny)
    a.tx, a.ty, a.coef = tabs["tx"].data_ptr(), tabs["ty"].data_ptr(), tabs["c"].data_ptr()
    a.ntx, a.nty = int(tabs["tx"].shape[0]), int(tabs["ty"].shape[0])
    a.w = 2 * np.pi * freq
    a.c = const.speed_of_light * 100
    a.n_crit = critical_den(freq, den_unit=getattr(n_field, "den_unit", "cm^-3"))
    a.x0, a.y0, a.kx0, a.ky0 = (t.data_ptr() for t in out)
    stream = torch.cuda.current_stream(dev)
    _cabi.check(_cabi.lib.stb_set_device(dev.index or 0))
    _cabi.check(_cabi.lib.stb_launch_states(ctypes.byref(a), ctypes.c_void_p(stream.cuda_stream)))
    return out[0], out[1], out[2], out[3], freq


def _host_results(res):
    """Per-ray outputs of a device-I/O BatchResult as numpy arrays (trajectories stay in HBM)."""
    for name in ("nsteps", "end_state", "turn_index", "turn_state", "lax_end", "ph_end", "cross_count", "cross_index",
                 "cross_flags", "cross_state", "lookup_xy"):
        v = getattr(res, name)
        if v is not None and hasattr(v, "cpu"):
            setattr(res, name, v.cpu().numpy())
    return res


def trace_rays(n_field, x0, y0, kx0, ky0, freq, amp=None, *, max_steps: Optional[int] = None,
               keep_traj: bool = True, want_extras: bool = False, tiled: bool = True,
               receiver=None, device_launch: Optional[bool] = None, host_results: bool = True) -> RayBundle:
    """Trace a whole fan of rays (replaces the serial loop of beam.py:340-342).

    max_steps: array length per ray, default ``Ray.MaxStep`` (ray.py:11) -- the reference's own cap, so a
    reflected or trapped ray runs exactly as long as ``Ray(...)`` lets it.  The cap costs no memory:
    a first pass without trajectory output yields every ray's step count, then the trajectories are
    written into an arena sized exactly for them (``sidtrace.trace_exact``: rays sorted by length into
    256-ray tiles; ``tiled=False``: ray-major [n, max(nsteps)+1, 4]).  ``keep_traj=False`` stops after
    the first pass (per-ray outputs only).  Rays that do hit the cap report nsteps == max_steps-1
    exactly like the reference.

    device_launch: prepare the launch states (ray.py:34-68) on the GPU (``launch_states_device``)
    instead of on the host; default: on for >= 65536 rays, where SciPy's point-wise spline
    evaluation would take longer than the trace.
    receiver: (val, axis) -- also report, per ray, what ``Ray.addpoint(val, axis)`` would
    (ray.py:131-224): first crossing of the receiver line, its phase and wave vector
    (``bundle.result.cross_*``), i.e. the per-ray inputs of ``Beam.field_r`` (beam.py:359-369).
    host_results: False leaves the per-ray outputs on the GPU (CUDA tensors in ``bundle.result`` /
    ``bundle.nsteps``) for the device-side consumers (``beam.field_result_device``); nothing O(rays) is
    copied to the host then."""
    import torch
    if device_launch is None:
        device_launch = np.size(x0) >= 65536 and hasattr(n_field, "device_launch_tables")
    if device_launch:
        d_in = list(launch_states_device(n_field, x0, y0, kx0, ky0, freq))
        freq = d_in.pop()
        xg, yg, kx, ky = (t.cpu().numpy() for t in d_in)
    else:
        xg, yg, kx, ky, freq = launch_states(n_field, x0, y0, kx0, ky0, freq)
        dev = torch.device("cuda", torch.cuda.current_device())
        d_in = [torch.from_numpy(a).to(dev) for a in (xg, yg, kx, ky)]
    dev = d_in[0].device
    n = len(xg)
    x0 = np.ascontiguousarray(np.atleast_1d(np.asarray(x0.cpu() if hasattr(x0, "cpu") else x0, np.float64)))
    y0 = np.ascontiguousarray(np.broadcast_to(np.asarray(y0.cpu() if hasattr(y0, "cpu") else y0, np.float64), x0.shape))
    if amp is None:
        amp = np.ones(n)
    amp = np.broadcast_to(np.asarray(amp, np.float64), (n,)).copy()
    L = int(max_steps) if max_steps else int(Ray.MaxStep)
    st = float(n_field.st * Ray.kStep)
    f = _device_field(n_field)
    kw = dict(L=L, st=st, dx=float(n_field.dx[0]), dy=float(n_field.dy[0]), want_extras=want_extras, cross=receiver,
              cells=_st.pick_cells(kx, ky))      # the cell table whose fast axis follows the fan (same cells, same bits)
    d_x0 = d_y0 = None
    if receiver or not host_results:
        d_x0, d_y0 = torch.from_numpy(x0.copy()).to(dev), torch.from_numpy(y0.copy()).to(dev)
    if receiver:
        kw.update(x0_phys=d_x0, y0_phys=d_y0)
    if keep_traj and tiled and n:
        res = _st.trace_exact(f, *d_in, **kw)
    else:
        res = _st.trace_batch(f, *d_in, **kw)          # O(rays) outputs; L only bounds the loop
        if keep_traj and n:
            # ray-major trajectories [n, L2, 4], L2 = states of the longest ray: the second pass takes
            # the same steps (no ray needs more than L2 - 1 <= L - 1)
            L2 = int(res.nsteps.max().item()) + 1
            free, _total = torch.cuda.mem_get_info(dev)
            need = n * L2 * 32
            if need > 0.9 * free:
                raise MemoryError(f"{n} rays x {L2} states x 32 B = {need / 1e9:.1f} GB of ray-major trajectories exceed the "
                                  f"free HBM ({free / 1e9:.1f} GB): use tiled=True (exactly sized), keep_traj=False, or chunks")
            t = _st.trace_batch(f, *d_in, L=L2, st=st, dx=kw["dx"], dy=kw["dy"], traj=True, want_end=False)
            res.traj, res.tiled = t.traj, False
    keep_dev = None
    if host_results:
        torch.cuda.current_stream(dev).synchronize()
        res = _host_results(res)
    else:
        keep_dev = {"launch": d_in, "x0": d_x0, "y0": d_y0, "amp": torch.from_numpy(amp).to(dev)}
    return RayBundle(n_field=n_field, x0=x0, y0=y0, freq=freq, st=st, amp=amp,
                     xg0=xg, yg0=yg, kx0=kx, ky0=ky, nsteps=res.nsteps, end_state=res.end_state, result=res, L=L,
                     receiver=(float(receiver[0]), receiver[1]) if receiver else None, dev=keep_dev)
