"""Beyond-reference physics mode: adaptive Runge-Kutta ray tracing through an
interpolated medium with O/X-mode cold-plasma dispersion, absorption and a deposition profile.

EXTENSION, NOT PARITY: the reference's only integrator is the fixed-step loop of
sidtracemodule.c:36-48 (``rayrtrace_b200.sidtrace``).  This module is the host side of
include/sidtrace_b200_rk.h; it is validated against analytic solutions, a SciPy integration and a
bit-level CPU checker in tests/test_rk_*.py, never against reference output.

    med = Medium.from_density(n, x, y, freq=136.268e9)          # X = n / n_c  (den.py:335-348)
    res = trace_rk(med, x0, y0, px0, py0)                       # CUDA tensors in, CUDA tensors out
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import numpy as np

from . import _cabi
from ._cabi import StbEquilibrium, StbMediumDesc, StbRk, lib
from .nfield import critical_den

LEFT_DOMAIN, TAU_MAX, MAX_STEPS, STEP_UNDERFLOW, START_OUTSIDE, STOP_LINE = range(6)
INTERP = {"bilinear": 0, "catmull-rom": 1, "bspline": 2}

ELECTRON_CYCLOTRON_HZ_PER_T = 2.799249e10


class Equilibrium:
    """A tabulated axisymmetric equilibrium with the fields of the reference's ``Equilibrium`` class
    (equilibrium.py:11-49, geqdsk-like; SI units): ``psi[nx, ny]`` on the rectangle ``[rleft, rleft + rdim] x
    [zmid - zdim/2, zmid + zdim/2]`` (R-major, "so (x,y) axis order is in geqdsk format"), ``fpol`` = R B_phi on a
    uniform psi grid from ``simagx`` to ``sibdry``.  Any object with these attributes will do -- an instance of
    the reference's own class included.  The library interpolates psi bilinearly, as the reference's
    ``_psi_interp = RectBivariateSpline(..., kx=1, ky=1)`` does (equilibrium.py:46-49)."""

    def __init__(self, psi, fpol, *, rleft, rdim, zmid, zdim, simagx, sibdry):
        self.psi = np.ascontiguousarray(psi, dtype=np.float64)
        self.fpol = np.ascontiguousarray(fpol, dtype=np.float64)
        self.nx, self.ny = self.psi.shape
        self.rleft, self.rdim, self.zmid, self.zdim = float(rleft), float(rdim), float(zmid), float(zdim)
        self.simagx, self.sibdry = float(simagx), float(sibdry)

    @classmethod
    def solovev(cls, *, R0=0.55, a=0.08, kappa=1.0, B0=2.2, psi_b=2.0e-3, nr=129, nz=129, half=0.1, nf=65):
        """Solov'ev equilibrium psi = psi_b/(a^2 R0^2) [R^2 Z^2/kappa^2 + (R^2 - R0^2)^2/4], F = B0 R0 (vacuum
        toroidal field), tabulated like a geqdsk file.  Defaults: the FT-2 size of equilibrium.py:16-29
        (R0 = 0.55 m, 0.2 m box, 2.2 T).  Analytic, so tests can compare B against the closed form."""
        R = np.linspace(R0 - half, R0 + half, nr)
        Z = np.linspace(-half, half, nz)
        RR, ZZ = np.meshgrid(R, Z, indexing="ij")
        psi = psi_b / (a * a * R0 * R0) * (RR ** 2 * ZZ ** 2 / kappa ** 2 + (RR ** 2 - R0 ** 2) ** 2 / 4)
        eq = cls(psi, np.full(nf, B0 * R0), rleft=R0 - half, rdim=2 * half, zmid=0.0, zdim=2 * half, simagx=0.0, sibdry=psi_b)
        eq.analytic = dict(R0=R0, a=a, kappa=kappa, B0=B0, psi_b=psi_b)
        return eq

    def analytic_B(self, R, Z):
        """(B_R, B_Z, B_phi) of the Solov'ev equilibrium at (R, Z) in metres (only for ``solovev`` instances)."""
        p = self.analytic
        c = p["psi_b"] / (p["a"] ** 2 * p["R0"] ** 2)
        return (-c * 2 * R * Z / p["kappa"] ** 2, c * (2 * Z ** 2 / p["kappa"] ** 2 + (R ** 2 - p["R0"] ** 2)),
                p["B0"] * p["R0"] / R)


def _equilibrium_struct(eq, freq, r0_cm, z0_cm):
    """stb_equilibrium_t for ``eq`` plus the arrays it points to (keep them alive while the struct is in use)."""
    psi = np.ascontiguousarray(eq.psi, dtype=np.float64)
    fpol = np.ascontiguousarray(eq.fpol, dtype=np.float64)
    if psi.ndim != 2 or min(psi.shape) < 2 or fpol.ndim != 1 or fpol.size < 2:
        raise ValueError("the equilibrium needs psi[nr >= 2, nz >= 2] and fpol[nf >= 2]")
    e = StbEquilibrium()
    e.struct_size = ctypes.sizeof(StbEquilibrium)
    e.nr, e.nz = psi.shape
    e.rleft, e.rdim, e.zmid, e.zdim = float(eq.rleft), float(eq.rdim), float(eq.zmid), float(eq.zdim)
    e.psi, e.fpol, e.nf = psi.ctypes.data, fpol.ctypes.data, fpol.size
    e.simagx, e.sibdry = float(eq.simagx), float(eq.sibdry)
    e.r0_cm, e.z0_cm = float(r0_cm), float(z0_cm)
    e.y_per_tesla = ELECTRON_CYCLOTRON_HZ_PER_T / float(freq)
    return e, (psi, fpol)


def _torch():
    import torch
    return torch


class Medium:
    """Device-resident medium: the n/n_c node table (and an optional absorption table) on a
    uniform grid, plus the dispersion mode.  ``x``/``y`` are the node coordinates (cm, uniform).
    interp: 'bspline' (C2 natural cubic spline, prefiltered on the device -- the smooth default),
    'catmull-rom' (C1 local bicubic patches) or 'bilinear'."""

    def __init__(self, X, x, y, *, alpha=None, mode: str = "O", Y0: float = 0.0, R0: float = 55.0,
                 interp: str = "bspline", Y2=None, equilibrium=None, freq: Optional[float] = None, z0: float = 0.0):
        """X mode: Y = Y0 R0/(R0 + x) (analytic toroidal field) unless ``Y2`` -- the [ny, nx] node table of
        (omega_ce/omega)^2 -- or ``equilibrium`` (an ``Equilibrium``; needs ``freq``) is given: then Y^2 is
        interpolated like X, from the caller's table or from |B| of the tabulated equilibrium, evaluated on the
        device at the nodes (R = (R0 + x)/100 m, Z = (y - z0)/100 m)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if X.ndim != 2 or X.shape != (y.size, x.size):
            raise ValueError(f"X must have shape (len(y), len(x)) = ({y.size}, {x.size}), got {X.shape}")
        if x.size < 2 or y.size < 2:
            raise ValueError("the table needs at least 2 x 2 nodes")
        dx, dy = (x[-1] - x[0]) / (x.size - 1), (y[-1] - y[0]) / (y.size - 1)
        if not (dx > 0 and dy > 0):
            raise ValueError("x and y must be increasing")
        if np.max(np.abs(np.diff(x) - dx)) > 1e-9 * dx or np.max(np.abs(np.diff(y) - dy)) > 1e-9 * dy:
            raise ValueError("the adaptive-RK medium needs a uniform grid")
        if mode not in ("O", "X"):
            raise ValueError("mode must be 'O' or 'X'")
        if interp not in INTERP:
            raise ValueError(f"interp must be one of {sorted(INTERP)}")
        if alpha is not None:
            alpha = np.ascontiguousarray(alpha, dtype=np.float64)
            if alpha.shape != X.shape:
                raise ValueError("alpha must have the shape of X")
        if (Y2 is not None or equilibrium is not None) and mode != "X":
            raise ValueError("Y2 / equilibrium belong to the X mode")
        if Y2 is not None and equilibrium is not None:
            raise ValueError("give Y2 or equilibrium, not both")
        if Y2 is not None:
            Y2 = np.ascontiguousarray(Y2, dtype=np.float64)
            if Y2.shape != X.shape:
                raise ValueError("Y2 must have the shape of X")
        if equilibrium is not None and not (freq and freq > 0):
            raise ValueError("equilibrium= needs freq= (Y = f_ce / freq)")
        self.tabulated_B = Y2 is not None or equilibrium is not None
        self.nx, self.ny = int(x.size), int(y.size)
        self.xorg, self.yorg, self.dx, self.dy = float(x[0]), float(y[0]), float(dx), float(dy)
        self.mode, self.Y0, self.R0, self.interp = mode, float(Y0), float(R0), interp
        self.has_alpha = alpha is not None
        torch = _torch()
        self.device = torch.device("cuda", torch.cuda.current_device())
        _cabi.check(lib.stb_set_device(self.device.index or 0))
        d = StbMediumDesc()
        d.struct_size = ctypes.sizeof(StbMediumDesc)
        d.nx, d.ny = self.nx, self.ny
        d.xorg, d.yorg, d.dx, d.dy = self.xorg, self.yorg, self.dx, self.dy
        d.X = X.ctypes.data
        d.alpha = alpha.ctypes.data if alpha is not None else None
        d.on_device = 0
        d.mode = (2 if self.tabulated_B else 1) if mode == "X" else 0
        d.interp = INTERP[interp]
        d.Y0, d.R0 = self.Y0, self.R0
        keep = None
        if Y2 is not None:
            d.Y2 = Y2.ctypes.data
        elif equilibrium is not None:
            es, keep = _equilibrium_struct(equilibrium, freq, self.R0, z0)
            d.equilibrium = ctypes.addressof(es)
        h = ctypes.c_void_p()
        _cabi.check(lib.stb_medium_create(ctypes.byref(d), ctypes.byref(h)))
        self._h = h

    @classmethod
    def from_density(cls, n, x, y, *, freq: float = 136.268e9, den_unit="cm^-3", alpha=None, mode: str = "O",
                     btor: float = 2.2, R0: float = 55.0, interp: str = "bspline", equilibrium=None, z0: float = 0.0) -> "Medium":
        """n in cm^-3 (times ``den_unit``) on nodes -> X = n / n_c(freq) (den.py:335-348); the
        X mode's Y0 = f_ce(btor) / freq with the rmain/btor defaults of equilibrium.py:289, or -- with
        ``equilibrium`` -- |B| from the tabulated psi(R, Z) and F(psi)."""
        X = np.asarray(n, dtype=np.float64) / critical_den(freq, den_unit=den_unit)
        Y0 = ELECTRON_CYCLOTRON_HZ_PER_T * btor / freq if mode == "X" else 0.0
        return cls(X, x, y, alpha=alpha, mode=mode, Y0=Y0, R0=R0, interp=interp, equilibrium=equilibrium, freq=freq, z0=z0)

    def table(self, which: str = "X") -> np.ndarray:
        """The table the kernel reads, copied back: node values, or for the B-spline the
        (ny+2, nx+2) coefficients the device prefilter produced."""
        pad = 2 if self.interp == "bspline" else 0
        out = np.empty((self.ny + pad, self.nx + pad))
        _cabi.check(lib.stb_medium_table(self._h, {"X": 0, "alpha": 1, "Y2": 2}[which], out.ctypes.data, out.size))
        return out

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.stb_medium_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class RkResult:
    __slots__ = ("end_state", "tau", "phi", "sig", "n2_min", "ham", "nacc", "nrej", "status", "turn_index",
                 "turn_state", "traj", "aux", "hist")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


def trace_rk(med: Medium, x0, y0, px0, py0, *, rtol: float = 1e-9, atol_r: float = 1e-9, atol_p: float = 1e-9,
             h0: Optional[float] = None, hmin: float = 1e-12, hmax: Optional[float] = None,
             tau_max: float = math.inf, max_steps: int = 100000, xscale=None, yscale=None, amp=None,
             diagnostics: bool = True, L: int = 0, aux: bool = False, hist=None, nbins: int = 0,
             xaxis: float = 0.0, yaxis: float = 0.0, r_lcs: float = 1.0, rho_max: float = 1.0,
             stop=None, stop_tol: float = 1e-12, refill_min: Optional[int] = None, lane_refill: bool = False,
             sort_rays: bool = False, out: Optional[RkResult] = None, stream=None) -> RkResult:
    """Trace ``len(x0)`` rays with the adaptive Dormand-Prince 5(4) kernel, one launch.

    x0, y0 (cm) and px0, py0 (= k/k0) are contiguous float64 CUDA tensors; so are the optional
    per-ray ``xscale`` / ``yscale`` (multipliers of X and Y: density and frequency sweeps) and
    ``amp``.  ``h0`` defaults to half a cell and ``hmax`` to four cells, which bounds how far the
    last step can carry a ray past the table edge.  ``L`` > 0 records the first L accepted states
    per ray ([nrays, L, 4], NaN-filled) and with ``aux`` their (tau, sigma).  ``nbins`` > 0 (or a
    preallocated ``hist``) accumulates absorbed power per rho bin; needs a medium with ``alpha``.
    ``stop=(val, 'y')`` (or ``'x'``) ends every ray ON the line y = val (a receiver line, the role of
    ``Ray.addpoint``, ray.py:131-224): the crossing step is repeated with shorter lengths until
    ``|y - val| <= stop_tol``; status STOP_LINE, ``res.phi`` is then the optical path to the line.
    Schedule: a warp claims new rays (consecutive indices) once ``refill_min`` of its lanes are idle.
    32 = warp tiles (fastest on a clean beam launched in order: neighbouring rays share cells and the
    warp's loads coalesce), 1 = ``lane_refill=True`` (random ray order, or a few rays taking 10-100x the
    steps of the rest); the default 24 is within 4 % of the tiles and robust against stragglers.
    ``sort_rays=True`` traces the rays in a coherent order (Morton order of launch cell and direction)
    and returns the results in the caller's order: for sweeps whose rays arrive in random order this
    restores the locality the kernel's loads depend on (about 2x); per-ray results are unchanged.
    Asynchronous on ``stream`` (default: the current torch stream).
    """
    torch = _torch()
    n = int(x0.shape[0])
    if sort_rays and n > 1:
        perm = coherent_order(med, x0, y0, px0, py0)
        pick = lambda a: None if a is None else a.index_select(0, perm).contiguous()
        sub = trace_rk(med, pick(x0), pick(y0), pick(px0), pick(py0), rtol=rtol, atol_r=atol_r, atol_p=atol_p, h0=h0,
                       hmin=hmin, hmax=hmax, tau_max=tau_max, max_steps=max_steps, xscale=pick(xscale), yscale=pick(yscale),
                       amp=pick(amp), diagnostics=diagnostics, L=L, aux=aux, hist=hist, nbins=nbins, xaxis=xaxis, yaxis=yaxis,
                       r_lcs=r_lcs, rho_max=rho_max, stop=stop, stop_tol=stop_tol, refill_min=refill_min,
                       lane_refill=lane_refill, sort_rays=False, stream=stream)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n, device=perm.device)
        res = out if out is not None else RkResult()
        for name in RkResult.__slots__:
            v = getattr(sub, name)
            if v is None or name == "hist":
                setattr(res, name, v if getattr(res, name) is None or name == "hist" else getattr(res, name))
                continue
            v = v.index_select(1 if name in ("end_state", "turn_state") else 0, inv)
            cur = getattr(res, name)
            if cur is not None and cur.shape == v.shape:
                cur.copy_(v)
            else:
                setattr(res, name, v)
        return res
    ins = {"x0": x0, "y0": y0, "px0": px0, "py0": py0}
    for name, a in (("xscale", xscale), ("yscale", yscale), ("amp", amp)):
        if a is not None:
            ins[name] = a
    for name, a in ins.items():
        if not isinstance(a, torch.Tensor) or a.dtype != torch.float64 or not a.is_cuda or not a.is_contiguous():
            raise TypeError(f"{name} must be a contiguous float64 CUDA tensor")
        if a.ndim != 1 or int(a.shape[0]) != n:
            raise ValueError(f"{name} must be 1-D of length {n}")
        if a.device != med.device:
            raise ValueError(f"{name} lives on {a.device} but the medium on {med.device}")
    cell = min(med.dx, med.dy)
    h0 = 0.5 * cell if h0 is None else float(h0)
    hmax = 4.0 * cell if hmax is None else float(hmax)
    res = out if out is not None else RkResult()
    dev = med.device

    def need(name, shape, dtype):
        cur = getattr(res, name)
        if cur is None:
            cur = torch.empty(shape, dtype=dtype, device=dev)
            setattr(res, name, cur)
        return cur

    need("end_state", (4, n), torch.float64)
    need("nacc", (n,), torch.int32)
    need("nrej", (n,), torch.int32)
    need("status", (n,), torch.int32)
    if diagnostics:
        for name in ("tau", "phi", "sig", "n2_min", "ham"):
            need(name, (n,), torch.float64)
        need("turn_index", (n,), torch.int32)
        need("turn_state", (4, n), torch.float64)
    if L:
        if res.traj is None:
            res.traj = torch.full((n, L, 4), math.nan, dtype=torch.float64, device=dev)
        if aux and res.aux is None:
            res.aux = torch.full((n, L, 2), math.nan, dtype=torch.float64, device=dev)
    if hist is None and nbins:
        hist = torch.zeros(nbins, dtype=torch.float64, device=dev)
    if hist is not None:
        if not med.has_alpha:
            raise ValueError("the deposition histogram needs a medium with an absorption table (alpha)")
        nbins = int(hist.shape[0])
        res.hist = hist

    k = StbRk()
    k.struct_size = ctypes.sizeof(StbRk)
    k.flags = 0
    k.refill_min = 1 if lane_refill else (0 if refill_min is None else int(refill_min))
    k.nrays = n
    k.x0, k.y0, k.px0, k.py0 = x0.data_ptr(), y0.data_ptr(), px0.data_ptr(), py0.data_ptr()
    k.xscale = xscale.data_ptr() if xscale is not None else None
    k.yscale = yscale.data_ptr() if yscale is not None else None
    k.amp = amp.data_ptr() if amp is not None else None
    k.rtol, k.atol_r, k.atol_p = float(rtol), float(atol_r), float(atol_p)
    k.h0, k.hmin, k.hmax, k.tau_max = h0, float(hmin), hmax, float(tau_max)
    k.max_steps = int(max_steps)
    if stop is not None:
        if stop[1] not in ("x", "y"):
            raise ValueError("stop axis must be 'x' or 'y'")
        k.stop_axis, k.stop_val, k.stop_tol = (1 if stop[1] == "y" else 2), float(stop[0]), float(stop_tol)
    k.end_state = res.end_state.data_ptr()
    k.nacc, k.nrej, k.status = res.nacc.data_ptr(), res.nrej.data_ptr(), res.status.data_ptr()
    if diagnostics:
        k.tau_end, k.phi_end, k.sig_end = res.tau.data_ptr(), res.phi.data_ptr(), res.sig.data_ptr()
        k.n2_min, k.ham_end = res.n2_min.data_ptr(), res.ham.data_ptr()
        k.turn_index, k.turn_state = res.turn_index.data_ptr(), res.turn_state.data_ptr()
    if L:
        k.traj = res.traj.data_ptr()
        k.traj_aux = res.aux.data_ptr() if res.aux is not None else None
        k.L = int(L)
    if hist is not None:
        k.hist, k.nbins = hist.data_ptr(), nbins
        k.xaxis, k.yaxis, k.r_lcs, k.rho_max = float(xaxis), float(yaxis), float(r_lcs), float(rho_max)
    if n == 0:
        return res
    _cabi.check(lib.stb_set_device(dev.index or 0))
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
    _cabi.check(lib.stb_trace_rk(med._h, ctypes.byref(k), sptr))
    return res


def coherent_order(med: Medium, x0, y0, px0, py0):
    """Permutation that puts rays with neighbouring launch cells and directions next to each other:
    30-bit Morton key of (x, y, direction angle), 10 bits each.  CUDA tensors in, int64 CUDA tensor out."""
    torch = _torch()

    def quant(t):
        t = torch.nan_to_num(t, nan=0.5, posinf=1.0, neginf=0.0).clamp(0.0, 1.0)
        return (t * 1023.0).to(torch.int64)

    def spread(q):      # 10 bits -> every third bit
        q = (q | (q << 16)) & 0x30000FF
        q = (q | (q << 8)) & 0x300F00F
        q = (q | (q << 4)) & 0x30C30C3
        return (q | (q << 2)) & 0x9249249

    u = (x0 - med.xorg) / (med.dx * (med.nx - 1))
    v = (y0 - med.yorg) / (med.dy * (med.ny - 1))
    a = (torch.atan2(py0, px0) + math.pi) / (2 * math.pi)
    key = spread(quant(u)) | (spread(quant(v)) << 1) | (spread(quant(a)) << 2)
    return torch.argsort(key)


def refractive_index(med: Medium, x0, y0, xscale=None, yscale=None):
    """N at the given points (CUDA tensors), evaluated by the kernel itself: a zero-step launch
    reports N^2 at the launch point.  Scale unit launch directions by it so that H = 0 (the role
    of ray.py:62-64 for the interpolated medium)."""
    torch = _torch()
    z = torch.zeros_like(x0)
    res = trace_rk(med, x0, y0, z, z, xscale=xscale, yscale=yscale, max_steps=0)
    return torch.sqrt(torch.clamp(res.n2_min, min=0.0))


def beam_to_receiver(beam, med: Medium, n_field, receiver, rnum=None, dist=None, ampbnd=None, **rk_kw):
    """A ``rayrtrace_b200.beam.Beam`` traced with the adaptive-RK integrator instead of the reference
    scheme: the same launch set as ``Beam.rays`` (beam.py:278-331: wavefront, equidistant start points,
    normals, amplitudes), each ray ended ON the receiver line ``receiver = (val, 'y' | 'x')``.

    Returns a dict of CUDA tensors: ``x``, ``y`` (where the ray meets the line), ``phase`` (eikonal
    phase from the launch point, rad: the quantity ``Ray.ph`` holds at the point ``Ray.addpoint``
    inserts, ray.py:113-224), ``amp`` (launch amplitude), ``status``.  The reference scheme's phases
    converge to these at first order in the grid spacing."""
    from scipy import constants as const
    torch = _torch()
    rnum = type(beam)._nray if rnum is None else rnum
    dist = type(beam)._dist if dist is None else dist
    ampbnd = type(beam)._ampbnd if ampbnd is None else ampbnd
    wave_front = beam.wavefront_r(dist=dist, ampbnd=ampbnd, n_field=n_field)
    n_edge = n_field.refractive_index(freq=beam.freq)
    x, y, kx, ky, amps, _ = beam.launch_conditions(rnum, wave_front, n_edge)
    scale = 100 if beam.length_unit == "m" else 1
    dev = med.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    x0, y0, dx, dy = t(x * scale), t(y * scale), t(kx), t(ky)
    n_loc = refractive_index(med, x0, y0)
    res = trace_rk(med, x0, y0, (n_loc * dx).contiguous(), (n_loc * dy).contiguous(), stop=receiver, **rk_kw)
    k0 = 2 * math.pi * beam.freq / (const.speed_of_light * 100)
    return {"x": res.end_state[0], "y": res.end_state[1], "phase": k0 * res.phi, "amp": t(amps), "status": res.status,
            "result": res}

