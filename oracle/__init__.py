"""CPU oracle for the sidtrace hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only
as the checker.  The product (``rayrtrace_b200``) never imports it.

Two arms:

* ``port``  -- ``libsidtrace_oracle.so``: our plain-C restatement
  (``sidtrace_oracle.c``), each function citing the reference file:line.
* ``ref``   -- ``oracle/_ref/sidtrace.<abi>.so``: the UNMODIFIED reference
  extension (``/root/reference/sidtracemodule.c``) compiled by
  ``oracle/Makefile`` where it lies.  Git-ignored; travels to the GPU box.

Parity status: PINNED (see sidtrace_oracle.c header and tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libsidtrace_oracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int32)


def build(ref: bool = True) -> None:
    """Compile the port (always) and the reference extension (when the
    reference sources are present -- i.e. in the build container)."""
    targets = ["all"]
    if ref and os.path.exists(os.path.join(REFERENCE_ROOT, "sidtracemodule.c")):
        targets.append("ref")
    subprocess.check_call(["make", "-s", "-C", _HERE] + targets)


_port = None


def port():
    """ctypes handle of the C restatement (built on demand)."""
    global _port
    if _port is None:
        if not os.path.exists(_PORT_SO):
            build(ref=False)
        lib = ctypes.CDLL(_PORT_SO)
        lib.oracle_trace.restype = ctypes.c_int64
        lib.oracle_trace.argtypes = [_c_dp] * 4 + [ctypes.c_int64] + [ctypes.c_double] * 3 + \
            [ctypes.c_int, ctypes.c_int, _c_dp, _c_dp]
        lib.oracle_check.restype = ctypes.c_double
        lib.oracle_check.argtypes = [ctypes.c_int] * 4 + [_c_dp]
        lib.oracle_trace_batch.restype = ctypes.c_int
        lib.oracle_trace_batch.argtypes = [ctypes.c_int64, ctypes.c_int64] + [_c_dp] * 4 + \
            [ctypes.c_double, _c_dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
             _c_dp, _c_dp, _c_dp, _c_ip, _c_dp, ctypes.c_int]
        lib.oracle_lax_ph.restype = ctypes.c_double
        lib.oracle_lax_ph.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double, _c_dp, _c_dp]
        lib.oracle_turning_point.restype = ctypes.c_int64
        lib.oracle_turning_point.argtypes = [_c_dp, ctypes.c_int64]
        lib.oracle_deposit.restype = None
        lib.oracle_deposit.argtypes = [_c_dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double] + \
            [ctypes.c_double] * 6 + [ctypes.c_int, _c_dp]
        _port = lib
    return _port


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_c_dp)


def trace(x, y, kx, ky, st, dx, dy, nx, ny, ax, ay) -> int:
    """Port of ``sidtrace.trace`` (sidtracemodule.c:8-51): same signature, in place."""
    for a in (x, y, kx, ky, ax, ay):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    return int(port().oracle_trace(_dp(x), _dp(y), _dp(kx), _dp(ky), len(x), float(st), float(dx),
                                   float(dy), int(nx), int(ny), _dp(ax), _dp(ay)))


def check(nx, ny, i, j, a) -> float:
    """Port of ``sidtrace.check`` (sidtracemodule.c:53-69)."""
    return float(port().oracle_check(int(nx), int(ny), int(i), int(j), _dp(a)))


def trace_batch(x0, y0, kx0, ky0, L, st, dx, dy, nx, ny, ax, ay, *, st_ray=None,
                want_traj=True, nthreads=1):
    """Trace a batch on the CPU.  Returns dict(nsteps, end[4,n], traj[n,L,4] or None).

    Output layout equals the CUDA product's so tests can compare bytes."""
    x0, y0, kx0, ky0 = (np.ascontiguousarray(v, np.float64) for v in (x0, y0, kx0, ky0))
    ax = np.ascontiguousarray(ax, np.float64)
    ay = np.ascontiguousarray(ay, np.float64)
    n = x0.shape[0]
    traj = np.zeros((n, L, 4), np.float64) if want_traj else None
    nsteps = np.zeros(n, np.int32)
    end = np.zeros((4, n), np.float64)
    if st_ray is not None:
        st_ray = np.ascontiguousarray(st_ray, np.float64)
    rc = port().oracle_trace_batch(n, int(L), _dp(x0), _dp(y0), _dp(kx0), _dp(ky0), float(st),
                                   _dp(st_ray), float(dx), float(dy), int(nx), int(ny), _dp(ax), _dp(ay),
                                   _dp(traj), nsteps.ctypes.data_as(_c_ip), _dp(end), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_trace_batch failed rc={rc}")
    return {"nsteps": nsteps, "end": end, "traj": traj}


def lax_ph(traj_kept, st):
    """ray.py:113-126 on a kept trajectory ``traj_kept[n,4]``; returns (lax, ph)."""
    t = np.ascontiguousarray(traj_kept, np.float64)
    n = t.shape[0]
    lax = np.zeros(n)
    ph = np.zeros(n)
    port().oracle_lax_ph(_dp(t), n, float(st), _dp(lax), _dp(ph))
    return lax, ph


def turning_point(traj, n) -> int:
    t = np.ascontiguousarray(traj, np.float64)
    return int(port().oracle_turning_point(_dp(t), int(n)))


def deposit(traj_kept, st, amp, xorg, yorg, dx, dy, r_lcs, rho_max, nbins, hist):
    t = np.ascontiguousarray(traj_kept, np.float64)
    port().oracle_deposit(_dp(t), t.shape[0], float(st), float(amp), float(xorg), float(yorg), float(dx),
                          float(dy), float(r_lcs), float(rho_max), int(nbins), _dp(hist))


# --------------------------------------------------------------------------
# The compiled, unmodified reference extension
# --------------------------------------------------------------------------
def ref_so_path() -> str:
    return os.path.join(_REF_DIR, "sidtrace" + sysconfig.get_config_var("EXT_SUFFIX"))


def have_ref() -> bool:
    return os.path.exists(ref_so_path())


_ref_mod = None


def ref():
    """Import the compiled reference extension as a module object (has
    ``trace`` and ``check`` exactly as sidtracemodule.c:71-75 exports them)."""
    global _ref_mod
    if _ref_mod is None:
        path = ref_so_path()
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` in the build container")
        loader = importlib.machinery.ExtensionFileLoader("sidtrace", path)
        spec = importlib.util.spec_from_loader("sidtrace", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _ref_mod = mod
    return _ref_mod
