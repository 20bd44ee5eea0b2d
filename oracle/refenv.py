"""Import the reference's own Python layers (den.py, ray.py, beam.py) in THIS
container, unmodified, on top of the compiled reference extension.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture
generation) and by tests that are skipped when /root/reference is absent (it
does not exist on the GPU box).

Shims (harness-side; /root/reference is never edited) -- SURVEY.md 8c:
  * scipy.interpolate.interp2d was removed in SciPy 1.14 (den.py:4,71 imports
    and calls it).  Replaced by a RectBivariateSpline(kx=3, ky=3, s=0) wrapper
    returning shape (1,) for scalar queries, like the old interp2d did.  Only
    the launch refractive index (ray.py:62) depends on it; that is an INPUT of
    the hot path, so parity tests feed the same k0 to both sides.
  * skimage (beam.py:5) is not installed: a stub module whose
    measure.find_contours restates the published marching-squares algorithm
    for the reference's one call site (beam.py:238; single open iso-line of a
    smooth field).  Everything around it is the reference's own code.
  * ufile (equilibrium.py:8) is not in the reference repo: empty stub.
"""
from __future__ import annotations

import sys
import types
import warnings

import numpy as np

from . import REFERENCE_ROOT, ref

_installed = False


class _Interp2dShim:
    def __init__(self, x, y, z, kind="linear", fill_value=None, **_):
        from scipy.interpolate import RectBivariateSpline
        k = {"linear": 1, "cubic": 3, "quintic": 5}[kind]
        self._s = RectBivariateSpline(np.asarray(x, float), np.asarray(y, float),
                                      np.asarray(z, float).T, kx=k, ky=k, s=0)

    def __call__(self, x, y):
        x = np.atleast_1d(np.asarray(x, float))
        y = np.atleast_1d(np.asarray(y, float))
        out = self._s(x, y).T  # (len(y), len(x)) like interp2d
        if out.shape[0] == 1:
            out = out[0]
        elif out.shape[1] == 1:
            out = out[:, 0]
        return out


def _find_contours(arr, level):
    """Restatement of ``skimage.measure.find_contours(image, level)`` (scikit-image, any release
    since 0.10; the reference pins no version and skimage is not installable here) for the one way
    the reference calls it (beam.py:238): the marching-squares iso-line of a smooth phase map that
    crosses the map from its left to its right wall and is single-valued in the column coordinate.

    Published algorithm: every cell edge whose two corner values straddle the level carries one
    contour vertex, placed by LINEAR interpolation of the level along that edge; vertices are
    chained into polylines in (row, col) array coordinates.  For a left-to-right open line that is a
    graph over the columns, chaining order = increasing column, which is also what the reference
    relies on (``contour[0][0, 1]`` is tested against the left wall, beam.py:239).
    Returns ``[ndarray (n, 2)]`` of (row, col) like skimage."""
    a = np.asarray(arr, float) - level
    pts = []
    nr, nc = a.shape
    # vertices on vertical edges (between rows r and r+1, integer column)
    for c in range(nc):
        col = a[:, c]
        for r in np.where(col[:-1] * col[1:] < 0)[0]:
            pts.append((r + col[r] / (col[r] - col[r + 1]), float(c)))
    # vertices on horizontal edges (between columns c and c+1, integer row)
    for r in range(nr):
        row = a[r, :]
        for c in np.where(row[:-1] * row[1:] < 0)[0]:
            pts.append((float(r), c + row[c] / (row[c] - row[c + 1])))
    pts = np.array(sorted(pts, key=lambda p: p[1]))
    return [pts]


def install(root=None, sidtrace_module=None):
    """`root`: where the unmodified reference sources lie (default /root/reference).
    `sidtrace_module`: what ``import sidtrace`` resolves to for the reference's ray.py:4 -- by default the
    compiled reference extension; tests/test_dropin_reference.py passes the B200 drop-in instead."""
    global _installed
    if _installed:
        return
    root = root or REFERENCE_ROOT
    import scipy.interpolate as si
    if not hasattr(si, "_orig_interp2d"):
        si._orig_interp2d = getattr(si, "interp2d", None)
        si.interp2d = _Interp2dShim
    sys.modules["sidtrace"] = sidtrace_module if sidtrace_module is not None else ref()
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        me = types.ModuleType("skimage.measure")
        me.find_contours = _find_contours
        sk.measure = me
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = me
    if "ufile" not in sys.modules:
        uf = types.ModuleType("ufile")
        uf.Ufile = type("Ufile", (), {})
        sys.modules["ufile"] = uf
    if root not in sys.path:
        sys.path.insert(0, root)
    _installed = True


def modules():
    """Returns (den, ray) reference modules, imported unmodified."""
    install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import den  # noqa
        import ray  # noqa
    return den, ray
