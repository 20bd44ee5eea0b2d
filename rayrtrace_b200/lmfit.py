"""MINPACK's lmdif for the three-parameter Gaussian fit of field_result (beam.py:463-465), driven by
reductions over the data instead of the data itself.

The reference calls ``scipy.optimize.curve_fit(f, x, |field|, p0)`` -- MINPACK ``lmdif`` (Levenberg-
Marquardt, Jacobian by forward differences, ftol = xtol = 1.49012e-8, gtol = 0, factor = 100,
maxfev = 800).  On the folded, many-page beams of an over-dense plasma |field| is not Gaussian, the cost
surface is flat, and WHERE lmdif stops depends on its trust-region path -- so to return the reference's
numbers the same algorithm is followed: same scaling (diag = running max of the Jacobian column norms),
same lmpar iteration for the Levenberg-Marquardt parameter, same gain-ratio updates of the step bound,
same stopping tests.  MINPACK factors J = QR; here only J^T J, J^T r and r^T r are available (one
reduction launch per trial point, ``stb_gauss_fit_sums``), and every quantity lmdif needs is a function of
those: the step solves (J^T J + par D^2) p = -J^T r, |J p|^2 = p^T J^T J p.  Published algorithm:
J. J. More, "The Levenberg-Marquardt algorithm: implementation and theory", 1978; MINPACK-1 lmdif / lmpar.
"""
from __future__ import annotations

import numpy as np

_EPSMCH = np.finfo(float).eps
_DWARF = np.finfo(float).tiny


def fd_steps(p):
    """fdjac2 with epsfcn = machine epsilon: h_j = sqrt(eps) * |p_j| (sqrt(eps) when p_j = 0)."""
    eps = np.sqrt(_EPSMCH)
    h = eps * np.abs(np.asarray(p, float))
    h[h == 0] = eps
    return h


def _lmpar(A, g, diag, delta, par):
    """MINPACK lmpar on the normal equations: the step p with |D p| ~ delta (within 10 %) and its
    Levenberg-Marquardt parameter.  Returns (par, p) with p solving (A + par D^2) p = -g."""
    n = len(g)
    D2 = np.diag(diag * diag)

    def solve(alpha):
        M = A + alpha * D2
        try:
            return np.linalg.solve(M, -g)
        except np.linalg.LinAlgError:
            return np.linalg.lstsq(M, -g, rcond=None)[0]

    # Gauss-Newton direction (least-squares solution when A is singular)
    try:
        x = np.linalg.solve(A, -g)
        singular = not np.all(np.isfinite(x))
    except np.linalg.LinAlgError:
        singular = True
    if singular:
        x = np.linalg.lstsq(A, -g, rcond=None)[0]
    dxnorm = np.linalg.norm(diag * x)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:
        return 0.0, x
    # lower bound parl (zero when the Jacobian is rank deficient)
    parl = 0.0
    if not singular:
        q = diag * (diag * x) / dxnorm
        try:
            t2 = float(q @ np.linalg.solve(A, q))
            if t2 > 0:
                parl = (fp / delta) / t2
        except np.linalg.LinAlgError:
            parl = 0.0
    gn = np.linalg.norm(g / diag)
    paru = gn / delta
    if paru == 0:
        paru = _DWARF / min(delta, 0.1)
    par = min(max(par, parl), paru)
    if par == 0:
        par = gn / dxnorm
    it = 0
    while True:
        it += 1
        if par == 0:
            par = max(_DWARF, 0.001 * paru)
        x = solve(par)
        dxnorm = np.linalg.norm(diag * x)
        temp = fp
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta or (parl == 0 and fp <= temp and temp < 0) or it == 10:
            break
        q = diag * (diag * x) / dxnorm
        t2 = float(q @ np.linalg.solve(A + par * D2, q))
        parc = (fp / delta) / t2
        if fp > 0:
            parl = max(parl, par)
        if fp < 0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return par, x


ILL_CONDITIONED = -1


def _well_conditioned(A, rcond=1e-9):
    """The normal equations lose what MINPACK's pivoted QR keeps once the Jacobian's columns are nearly
    dependent (few rays inside the fitted Gaussian): the smallest eigenvalue of the column-scaled
    J^T J must stand clear of its rounding noise."""
    d = np.sqrt(np.maximum(np.diag(A), 0.0))
    if np.any(d == 0) or not np.all(np.isfinite(A)):
        return False
    w = np.linalg.eigvalsh(A / np.outer(d, d))
    return w[0] > rcond * w[-1]


def lmdif3(sums, p0, ftol=1.49012e-8, xtol=1.49012e-8, gtol=0.0, maxfev=None, factor=100.0):
    """``sums(p, h)`` -> (A = J^T J [n, n], g = J^T r [n], rr = r^T r) at parameters p with forward-difference
    steps h.  Returns (p, info) with MINPACK's info codes (1-4: converged, 5: maxfev, ...), or
    ``ILL_CONDITIONED`` as soon as a Jacobian is too close to rank deficient for the normal equations --
    the caller then runs the QR-based routine on the data itself."""
    x = np.array(p0, float)
    n = len(x)
    if maxfev is None:
        maxfev = 200 * (n + 1)
    nfev = 1
    A, g, rr = sums(x, fd_steps(x))
    nfev += n
    fnorm = np.sqrt(rr)
    par = 0.0
    it = 1
    diag = np.ones(n)
    delta = xnorm = 0.0
    info = 0
    while True:
        if not _well_conditioned(A):
            return x, ILL_CONDITIONED
        acnorm = np.sqrt(np.maximum(np.diag(A), 0.0))
        if it == 1:
            diag = np.where(acnorm == 0, 1.0, acnorm)
            xnorm = np.linalg.norm(diag * x)
            delta = factor * xnorm
            if delta == 0:
                delta = factor
        gnorm = 0.0
        if fnorm != 0:
            for j in range(n):
                if acnorm[j] != 0:
                    gnorm = max(gnorm, abs(g[j] / fnorm) / acnorm[j])
        if gnorm <= gtol:
            info = 4
            break
        diag = np.maximum(diag, acnorm)
        while True:
            par, p = _lmpar(A, g, diag, delta, par)
            x2 = x + p
            pnorm = np.linalg.norm(diag * p)
            if it == 1:
                delta = min(delta, pnorm)
            A2, g2, rr2 = sums(x2, fd_steps(x2))
            nfev += 1
            fnorm1 = np.sqrt(rr2) if np.isfinite(rr2) else np.inf
            actred = -1.0
            if 0.1 * fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            temp1 = np.sqrt(max(float(p @ A @ p), 0.0)) / fnorm
            temp2 = np.sqrt(par) * pnorm / fnorm
            prered = temp1 ** 2 + temp2 ** 2 / 0.5
            dirder = -(temp1 ** 2 + temp2 ** 2)
            ratio = actred / prered if prered != 0 else 0.0
            if ratio <= 0.25:
                temp = 0.5 if actred >= 0 else 0.5 * dirder / (dirder + 0.5 * actred)
                if 0.1 * fnorm1 >= fnorm or temp < 0.1:
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par = par / temp
            elif par == 0 or ratio >= 0.75:
                delta = pnorm / 0.5
                par = 0.5 * par
            if ratio >= 1e-4:
                x, A, g, rr = x2, A2, g2, rr2
                nfev += n                       # the Jacobian lmdif would now difference at the new point
                xnorm = np.linalg.norm(diag * x)
                fnorm = fnorm1
                it += 1
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1:
                info = 1
            if delta <= xtol * xnorm:
                info = 2
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1 and info == 2:
                info = 3
            if info != 0:
                break
            if nfev >= maxfev:
                info = 5
            if abs(actred) <= _EPSMCH and prered <= _EPSMCH and 0.5 * ratio <= 1:
                info = 6
            if delta <= _EPSMCH * xnorm:
                info = 7
            if gnorm <= _EPSMCH:
                info = 8
            if info != 0:
                break
            if ratio >= 1e-4:
                break
        if info != 0:
            break
    return x, info
