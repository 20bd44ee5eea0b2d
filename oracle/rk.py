"""ctypes wrapper of oracle/rk_oracle.c -- TEST INFRASTRUCTURE ONLY, PARITY UNPINNED.

The checker of the adaptive-RK extension (include/sidtrace_b200_rk.h).  The reference has no
counterpart; see the header of rk_oracle.c.  Never imported by ``rayrtrace_b200``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librk_oracle.so")
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)

LEFT_DOMAIN, TAU_MAX, MAX_STEPS, STEP_UNDERFLOW, START_OUTSIDE, STOP_LINE = range(6)


class _Medium(ctypes.Structure):
    _fields_ = [("X", _dp), ("alpha", _dp), ("nx", ctypes.c_int), ("ny", ctypes.c_int),
                ("xorg", ctypes.c_double), ("yorg", ctypes.c_double), ("rdx", ctypes.c_double), ("rdy", ctypes.c_double),
                ("Y0R0", ctypes.c_double), ("R0", ctypes.c_double), ("xmode", ctypes.c_int), ("interp", ctypes.c_int),
                ("Y2", _dp)]


class _Ctl(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("atol_r", ctypes.c_double), ("atol_p", ctypes.c_double),
                ("h0", ctypes.c_double), ("hmin", ctypes.c_double), ("hmax", ctypes.c_double), ("tau_max", ctypes.c_double),
                ("max_steps", ctypes.c_int), ("stop_axis", ctypes.c_int), ("stop_val", ctypes.c_double),
                ("stop_tol", ctypes.c_double), ("hist", _dp), ("nbins", ctypes.c_int),
                ("xaxis", ctypes.c_double), ("yaxis", ctypes.c_double), ("r_rlcs", ctypes.c_double),
                ("bins_per_rho", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
        L = ctypes.CDLL(_SO)
        L.rk_oracle_eval.restype = None
        L.rk_oracle_eval.argtypes = [ctypes.POINTER(_Medium)] + [ctypes.c_double] * 4 + [_dp] * 4
        L.rk_oracle_prefilter.restype = None
        L.rk_oracle_prefilter.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _dp]
        L.rk_oracle_equilibrium_y2.restype = None
        L.rk_oracle_equilibrium_y2.argtypes = [_dp, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4 + \
            [_dp, ctypes.c_int] + [ctypes.c_double] * 5 + [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4 + [_dp]
        L.rk_oracle_batch.restype = ctypes.c_int
        L.rk_oracle_batch.argtypes = [ctypes.POINTER(_Medium), ctypes.POINTER(_Ctl), ctypes.c_longlong] + [_dp] * 7 + \
            [_dp] * 6 + [_ip] * 3 + [_dp, _dp, ctypes.c_longlong, ctypes.c_int, _ip, _dp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a, n=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        a = np.ascontiguousarray(np.broadcast_to(a, (n,)))
    return a


BILINEAR, CATMULL_ROM, BSPLINE = 0, 1, 2


def prefilter(nodes):
    """[ny, nx] node values -> [ny+2, nx+2] natural cubic B-spline coefficients."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    ny, nx = nodes.shape
    coef = np.empty((ny + 2, nx + 2))
    lib().rk_oracle_prefilter(_p(nodes), nx, ny, _p(coef))
    return coef


def equilibrium_y2(eq, nx, ny, xorg, yorg, dx, dy):
    """[ny, nx] node table of Y^2 = (omega_ce/omega)^2 from a tabulated equilibrium: ``eq`` carries the fields of
    stb_equilibrium_t (psi [nr, nz], rleft, rdim, zmid, zdim, fpol, simagx, sibdry, r0_cm, z0_cm, y_per_tesla)."""
    psi = np.ascontiguousarray(eq["psi"], dtype=np.float64)
    fpol = np.ascontiguousarray(eq["fpol"], dtype=np.float64)
    out = np.empty((ny, nx))
    lib().rk_oracle_equilibrium_y2(_p(psi), psi.shape[0], psi.shape[1], float(eq["rleft"]), float(eq["rdim"]), float(eq["zmid"]),
                                   float(eq["zdim"]), _p(fpol), fpol.size, float(eq["simagx"]), float(eq["sibdry"]),
                                   float(eq["r0_cm"]), float(eq["z0_cm"]), float(eq["y_per_tesla"]), int(nx), int(ny),
                                   float(xorg), float(yorg), float(dx), float(dy), _p(out))
    return out


class Medium:
    """Host copy of the medium the kernel reads (stb_medium_desc_t)."""

    def __init__(self, X, xorg, yorg, dx, dy, *, alpha=None, mode="O", Y0=0.0, R0=55.0, interp=BSPLINE, Y2=None):
        """X, alpha, Y2: [ny, nx] NODE values; for interp = BSPLINE they are prefiltered here.
        mode "XT" = X mode with Y^2 from the node table Y2 (STB_RK_MODE_X_TABLE)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        ny, nx = X.shape
        if mode == "XT" and Y2 is None:
            raise ValueError("mode 'XT' needs the Y2 node table")
        if interp == BSPLINE:
            X = prefilter(X)
            alpha = None if alpha is None else prefilter(alpha)
            Y2 = None if Y2 is None else prefilter(Y2)
        self.X = X
        self.alpha = None if alpha is None else np.ascontiguousarray(alpha, dtype=np.float64)
        self.Y2 = None if Y2 is None else np.ascontiguousarray(Y2, dtype=np.float64)
        self.c = _Medium(_p(self.X), _p(self.alpha), nx, ny, float(xorg), float(yorg), 1.0 / float(dx), 1.0 / float(dy),
                         float(Y0) * float(R0), float(R0), {"O": 0, "X": 1, "XT": 2}[mode], int(interp), _p(self.Y2))

    def eval(self, x, y, xs=1.0, ys=1.0):
        """(fx, fy, alpha, N^2) at one point."""
        out = [ctypes.c_double() for _ in range(4)]
        lib().rk_oracle_eval(ctypes.byref(self.c), xs, ys, float(x), float(y), *[ctypes.byref(o) for o in out])
        return tuple(o.value for o in out)


def trace(med: Medium, x0, y0, px0, py0, *, rtol=1e-9, atol_r=1e-9, atol_p=1e-9, h0=0.05, hmin=1e-12, hmax=1e30,
          tau_max=float("inf"), max_steps=100000, xscale=None, yscale=None, amp=None, nbins=0, xaxis=0.0, yaxis=0.0,
          r_lcs=1.0, rho_max=1.0, L=0, aux=False, threads=1, stop=None, stop_tol=1e-12):
    """stop = (val, 'y' | 'x'): end the ray on that line."""
    x0 = _f64(np.atleast_1d(x0))
    n = x0.shape[0]
    y0, px0, py0 = _f64(y0, n), _f64(px0, n), _f64(py0, n)
    xscale, yscale, amp = _f64(xscale, n), _f64(yscale, n), _f64(amp, n)
    hist = np.zeros(nbins) if nbins else None
    stop_axis = 0 if stop is None else (1 if stop[1] == "y" else 2)
    ctl = _Ctl(rtol, atol_r, atol_p, h0, hmin, hmax, tau_max, int(max_steps), stop_axis,
               0.0 if stop is None else float(stop[0]), float(stop_tol), _p(hist), int(nbins), xaxis, yaxis,
               1.0 / r_lcs, (nbins / rho_max) if nbins else 0.0)
    end = np.empty((4, n))
    tau, phi, sig, n2min, ham = (np.empty(n) for _ in range(5))
    nacc, nrej, status, turn = (np.empty(n, dtype=np.int32) for _ in range(4))
    turn_state = np.empty((4, n))
    traj = np.full((n, L, 4), np.nan) if L else None
    auxa = np.full((n, L, 2), np.nan) if (L and aux) else None
    lib().rk_oracle_batch(ctypes.byref(med.c), ctypes.byref(ctl), n, _p(x0), _p(y0), _p(px0), _p(py0), _p(xscale),
                          _p(yscale), _p(amp), _p(end), _p(tau), _p(phi), _p(sig), _p(n2min), _p(ham), _pi(nacc), _pi(nrej),
                          _pi(status), _p(traj), _p(auxa), int(L), int(threads), _pi(turn), _p(turn_state))
    return dict(end_state=end, tau=tau, phi=phi, sig=sig, n2_min=n2min, ham=ham, nacc=nacc, nrej=nrej, status=status,
                turn_index=turn, turn_state=turn_state, traj=traj, aux=auxa, hist=hist)
