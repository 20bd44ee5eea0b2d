"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/sidtrace_b200.h declares; the ctypes mirror of stb_batch_t matches the
C layout.  No compute calls here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sidtrace_b200.h")
HEADER_RK = os.path.join(ROOT, "include", "sidtrace_b200_rk.h")   # adaptive-RK extension
HEADER_COMM = os.path.join(ROOT, "include", "sidtrace_b200_comm.h")   # multi-GPU entry points


def _declared():
    names = set()
    for h in (HEADER, HEADER_RK, HEADER_COMM):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(stb_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_expected_symbols():
    from rayrtrace_b200 import _cabi
    assert _declared() == sorted(_cabi.EXPORTS)


def test_library_exports_every_declared_symbol():
    from rayrtrace_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} not exported by {_cabi.LIB_PATH}"
    assert _cabi.lib.stb_abi_version() == _cabi.STB_ABI_VERSION
    assert _cabi.lib.stb_last_error() is not None


@pytest.mark.parametrize("cname, pyname", [("stb_batch_t", "StbBatch"), ("stb_launch_t", "StbLaunch"),
                                           ("stb_beam_t", "StbBeam"), ("stb_tube_t", "StbTube"), ("stb_overlap_t", "StbOverlap"),
                                           ("stb_medium_desc_t", "StbMediumDesc"), ("stb_equilibrium_t", "StbEquilibrium"),
                                           ("stb_rk_t", "StbRk")])
def test_struct_layout_matches_c(tmp_path, cname, pyname):
    from rayrtrace_b200 import _cabi
    cls = getattr(_cabi, pyname)
    fields = [f[0] for f in cls._fields_]
    prog = tmp_path / "layout.c"
    body = "\n".join(f'    printf("{f} %zu\\n", offsetof({cname}, {f}));' for f in fields)
    prog.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER_RK}"\n'
                    f'int main(void) {{\n    printf("sizeof %zu\\n", sizeof({cname}));\n{body}\n    return 0;\n}}\n')
    exe = tmp_path / "layout"
    subprocess.check_call([os.environ.get("CC", "gcc"), str(prog), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(cls)
    for f in fields:
        assert int(out[f]) == getattr(cls, f).offset, f


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback: importing the product without the CUDA library raises."""
    env = dict(os.environ, SIDTRACE_B200_LIB=str(tmp_path / "nope.so"), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", "import rayrtrace_b200.sidtrace"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr.replace("There is no CPU fallback", "no CPU fallback")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rayrtrace_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "sidtrace_oracle" not in text, f


def test_compute_entry_points_fail_loudly_without_a_device():
    """No CPU fallback: on a box without a GPU the field upload returns STB_ERR_CUDA with a message
    (and the Python surface raises) instead of computing anything on the host."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from rayrtrace_b200 import _cabi, sidtrace
    a = np.zeros((4, 6))
    h = ctypes.c_void_p()
    rc = _cabi.lib.stb_field_create(a.ctypes.data, a.ctypes.data, 7, 5, 0, ctypes.byref(h))
    assert rc == _cabi.STB_ERR_CUDA and not h.value
    assert b"cuda" in _cabi.lib.stb_last_error().lower()
    x = [np.zeros(8) for _ in range(4)]
    with pytest.raises(_cabi.SidtraceError):
        sidtrace.trace(*x, 0.1, 1.0, 1.0, 7, 5, a, a.copy())
    # the adaptive-RK extension has no host path either
    from rayrtrace_b200 import physics
    d = _cabi.StbMediumDesc()
    d.struct_size = ctypes.sizeof(_cabi.StbMediumDesc)
    X = np.zeros((4, 4))
    d.nx = d.ny = 4
    d.dx = d.dy = 1.0
    d.X = X.ctypes.data
    d.interp = 2
    m = ctypes.c_void_p()
    assert _cabi.lib.stb_medium_create(ctypes.byref(d), ctypes.byref(m)) == _cabi.STB_ERR_CUDA and not m.value
    with pytest.raises(Exception):
        physics.Medium(X, np.arange(4.0), np.arange(4.0))

