"""ctypes binding of the C ABI declared in include/sidtrace_b200.h.

The CUDA library is the product: if it is missing this module raises at import
time -- there is no CPU fallback anywhere in ``rayrtrace_b200``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t,
                    c_uint32, c_void_p)

from . import build as _build

STB_ABI_VERSION = 3
STB_OK = 0
STB_ERR_ARG, STB_ERR_CUDA, STB_ERR_NOMEM, STB_ERR_RANGE, STB_ERR_COMM = -1, -2, -3, -4, -5
STB_COMM_ID_BYTES = 128

F_HOST_IO = 1 << 0
F_DIV_IEEE = 1 << 1
F_REFILL = 1 << 2
F_TRAJ_HOST = 1 << 3
F_TRAJ_TILED = 1 << 4
F_STATIC = 1 << 5
F_NO_HOIST = 1 << 6
F_YMAJOR = 1 << 7

# every symbol include/sidtrace_b200.h declares (tests check the .so exports them all)
EXPORTS = (
    "stb_abi_version", "stb_last_error", "stb_device_count", "stb_set_device",
    "stb_field_create", "stb_field_destroy", "stb_field_shape", "stb_field_peek",
    "stb_trace", "stb_check", "stb_trace_batch", "stb_launch_states", "stb_beam_launch", "stb_field_pages", "stb_field_pages_links", "stb_tube_field", "stb_hole_fill", "stb_overlap", "stb_gauss_fit_sums", "stb_pack_traj", "stb_pack_traj_plan", "stb_traj_plan", "stb_launch_count",
    "stb_measure_fp64_peak", "stb_measure_store_bw", "stb_measure_store_pattern",
    # include/sidtrace_b200_comm.h (multi-GPU entry points)
    "stb_comm_unique_id", "stb_comm_init", "stb_comm_destroy", "stb_comm_rank", "stb_comm_allreduce_f64",
    "stb_comm_allreduce_max_f64", "stb_comm_gather", "stb_comm_allgather", "stb_shard_local_blocks", "stb_shard_local_rays",
    "stb_unshuffle_blocks",
    # include/sidtrace_b200_rk.h (adaptive-RK extension)
    "stb_medium_create", "stb_medium_destroy", "stb_medium_table", "stb_trace_rk",
)


class StbBatch(Structure):
    """Mirror of ``stb_batch_t``."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("flags", c_uint32),
        ("nrays", c_int64),
        ("L", c_int64),
        ("x0", c_void_p), ("y0", c_void_p), ("kx0", c_void_p), ("ky0", c_void_p),
        ("st", c_double),
        ("st_ray", c_void_p),
        ("dx", c_double), ("dy", c_double),
        ("traj", c_void_p),
        ("traj_ray_stride", c_int64),
        ("traj_tile", c_int32),
        ("nsteps", c_void_p),
        ("end_state", c_void_p),
        ("turn_index", c_void_p),
        ("turn_state", c_void_p),
        ("lax_end", c_void_p),
        ("ph_end", c_void_p),
        ("hist", c_void_p),
        ("nbins", c_int32),
        ("amp", c_void_p),
        ("xorg", c_double), ("yorg", c_double),
        ("r_lcs", c_double), ("rho_max", c_double),
        ("cross_axis", c_int32),
        ("cross_val", c_double),
        ("x0_phys", c_void_p), ("y0_phys", c_void_p),
        ("cross_count", c_void_p), ("cross_index", c_void_p), ("cross_flags", c_void_p),
        ("cross_state", c_void_p),
        ("lookup_ph", c_void_p), ("lookup_xy", c_void_p),
        ("traj_perm", c_void_p), ("traj_tile_off", c_void_p), ("traj_nslots", c_int64),
    ]


class StbLaunch(Structure):
    """Mirror of ``stb_launch_t``."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("n", c_int64),
        ("x0_phys", c_void_p), ("y0_phys", c_void_p),
        ("kx_dir", c_void_p), ("ky_dir", c_void_p),
        ("x_nodes", c_void_p), ("y_nodes", c_void_p),
        ("nx", c_int32), ("ny", c_int32),
        ("tx", c_void_p), ("ty", c_void_p), ("coef", c_void_p),
        ("ntx", c_int32), ("nty", c_int32),
        ("w", c_double), ("c", c_double), ("n_crit", c_double),
        ("x0", c_void_p), ("y0", c_void_p), ("kx0", c_void_p), ("ky0", c_void_p),
    ]


class StbBeam(Structure):
    """Mirror of ``stb_beam_t``."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("n", c_int64),
        ("nfront", c_int32),
        ("lax", c_void_p), ("wf_x", c_void_p),
        ("nw", c_int32),
        ("t", c_void_p), ("c", c_void_p),
        ("nt", c_int32),
        ("x0", c_double), ("y0", c_double), ("kx0", c_double), ("ky0", c_double), ("focus", c_double),
        ("waist", c_double), ("rayleigh", c_double),
        ("x", c_void_p), ("y", c_void_p), ("kx", c_void_p), ("ky", c_void_p), ("amp", c_void_p),
    ]


class StbTube(Structure):
    """Mirror of ``stb_tube_t``."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("n", c_int64),
        ("pred", c_void_p), ("succ", c_void_p), ("page_size", c_void_p),
        ("x0", c_void_p), ("y0", c_void_p),
        ("xc", c_void_p), ("yc", c_void_p), ("ph", c_void_p),
        ("amp0", c_void_p),
        ("lookup_xy", c_void_p),
        ("amp1", c_void_p),
        ("field", c_void_p),
    ]


class StbOverlap(Structure):
    """Mirror of ``stb_overlap_t``."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("n", c_int64),
        ("x", c_void_p),
        ("field_tx", c_void_p),
        ("y_rx", c_double),
        ("bx0", c_double), ("by0", c_double), ("bkx", c_double), ("bky", c_double), ("focus", c_double),
        ("waist", c_double), ("wave_len", c_double), ("rayleigh", c_double), ("gouy", c_double),
        ("field", c_void_p),
        ("absf", c_void_p),
        ("sums", c_void_p),
    ]


class StbMediumDesc(Structure):
    """Mirror of ``stb_medium_desc_t`` (include/sidtrace_b200_rk.h)."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("nx", c_int32), ("ny", c_int32),
        ("xorg", c_double), ("yorg", c_double),
        ("dx", c_double), ("dy", c_double),
        ("X", c_void_p), ("alpha", c_void_p),
        ("on_device", c_int32), ("mode", c_int32), ("interp", c_int32),
        ("Y0", c_double), ("R0", c_double),
        ("Y2", c_void_p), ("equilibrium", c_void_p),
    ]


class StbEquilibrium(Structure):
    """Mirror of ``stb_equilibrium_t`` (include/sidtrace_b200_rk.h)."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("nr", c_int32), ("nz", c_int32),
        ("rleft", c_double), ("rdim", c_double), ("zmid", c_double), ("zdim", c_double),
        ("psi", c_void_p), ("fpol", c_void_p),
        ("nf", c_int32),
        ("simagx", c_double), ("sibdry", c_double),
        ("r0_cm", c_double), ("z0_cm", c_double),
        ("y_per_tesla", c_double),
    ]


class StbRk(Structure):
    """Mirror of ``stb_rk_t`` (include/sidtrace_b200_rk.h)."""
    _fields_ = [
        ("struct_size", c_size_t),
        ("flags", c_uint32),
        ("nrays", c_int64),
        ("x0", c_void_p), ("y0", c_void_p), ("px0", c_void_p), ("py0", c_void_p),
        ("xscale", c_void_p), ("yscale", c_void_p), ("amp", c_void_p),
        ("rtol", c_double), ("atol_r", c_double), ("atol_p", c_double),
        ("h0", c_double), ("hmin", c_double), ("hmax", c_double),
        ("tau_max", c_double),
        ("max_steps", c_int32),
        ("refill_min", c_int32),
        ("stop_axis", c_int32),
        ("stop_val", c_double), ("stop_tol", c_double),
        ("end_state", c_void_p),
        ("tau_end", c_void_p), ("phi_end", c_void_p), ("sig_end", c_void_p), ("n2_min", c_void_p), ("ham_end", c_void_p),
        ("nacc", c_void_p), ("nrej", c_void_p), ("status", c_void_p),
        ("turn_index", c_void_p), ("turn_state", c_void_p),
        ("traj", c_void_p), ("traj_aux", c_void_p),
        ("L", c_int64),
        ("hist", c_void_p),
        ("nbins", c_int32),
        ("xaxis", c_double), ("yaxis", c_double),
        ("r_lcs", c_double), ("rho_max", c_double),
    ]


class SidtraceError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sidtrace_b200 error {code}: {msg}")
        self.code = code


def _load() -> ctypes.CDLL:
    path = os.environ.get("SIDTRACE_B200_LIB", _build.LIB)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the sidtrace_b200 CUDA library is not built. "
            "Run `python -m rayrtrace_b200.build` (needs nvcc). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    lib.stb_abi_version.restype = c_int
    lib.stb_abi_version.argtypes = []
    lib.stb_last_error.restype = c_char_p
    lib.stb_last_error.argtypes = []
    lib.stb_device_count.argtypes = [POINTER(c_int)]
    lib.stb_set_device.argtypes = [c_int]
    lib.stb_field_create.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]
    lib.stb_field_destroy.argtypes = [c_void_p]
    lib.stb_field_shape.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int)]
    lib.stb_field_peek.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_double)]
    lib.stb_trace.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_double, c_double, c_double, POINTER(c_int64)]
    lib.stb_check.argtypes = [c_int, c_int, c_int, c_int, c_void_p, POINTER(c_double)]
    lib.stb_trace_batch.argtypes = [c_void_p, POINTER(StbBatch), c_void_p]
    lib.stb_launch_states.argtypes = [POINTER(StbLaunch), c_void_p]
    lib.stb_launch_states.restype = c_int
    lib.stb_beam_launch.argtypes = [POINTER(StbBeam), c_void_p]
    lib.stb_beam_launch.restype = c_int
    lib.stb_field_pages.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, POINTER(c_int64)]
    lib.stb_field_pages.restype = c_int
    lib.stb_field_pages_links.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, POINTER(c_int64)]
    lib.stb_field_pages_links.restype = c_int
    lib.stb_tube_field.argtypes = [POINTER(StbTube), c_void_p]
    lib.stb_tube_field.restype = c_int
    lib.stb_hole_fill.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.stb_hole_fill.restype = c_int
    lib.stb_overlap.argtypes = [POINTER(StbOverlap), c_void_p]
    lib.stb_overlap.restype = c_int
    lib.stb_gauss_fit_sums.argtypes = [c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p]
    lib.stb_gauss_fit_sums.restype = c_int
    lib.stb_pack_traj.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.stb_pack_traj.restype = c_int
    lib.stb_pack_traj_plan.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.stb_pack_traj_plan.restype = c_int
    lib.stb_traj_plan.argtypes = [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_void_p]
    lib.stb_traj_plan.restype = c_int
    lib.stb_launch_count.restype = c_int64
    lib.stb_launch_count.argtypes = []
    lib.stb_measure_fp64_peak.argtypes = [POINTER(c_double)]
    lib.stb_measure_store_bw.argtypes = [c_size_t, POINTER(c_double)]
    lib.stb_measure_store_pattern.argtypes = [c_int64, c_int32, c_int64, c_int32, c_int32, POINTER(c_double)]
    lib.stb_measure_store_pattern.restype = c_int
    lib.stb_comm_unique_id.argtypes = [c_void_p]
    lib.stb_comm_init.argtypes = [c_void_p, c_int, c_int, POINTER(c_void_p)]
    lib.stb_comm_destroy.argtypes = [c_void_p]
    lib.stb_comm_rank.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int)]
    lib.stb_comm_allreduce_f64.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.stb_comm_allreduce_max_f64.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.stb_comm_gather.argtypes = [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]
    lib.stb_comm_allgather.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    lib.stb_shard_local_blocks.argtypes = [c_int64, c_int32, c_int32]
    lib.stb_shard_local_blocks.restype = c_int64
    lib.stb_shard_local_rays.argtypes = [c_int64, c_int32, c_int32, c_int32]
    lib.stb_shard_local_rays.restype = c_int64
    lib.stb_unshuffle_blocks.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32, c_int32, c_int64, c_int64, c_int64, c_void_p]
    for name in ("stb_comm_unique_id", "stb_comm_init", "stb_comm_destroy", "stb_comm_rank", "stb_comm_allreduce_f64",
                 "stb_comm_allreduce_max_f64", "stb_comm_gather", "stb_comm_allgather", "stb_unshuffle_blocks"):
        getattr(lib, name).restype = c_int
    lib.stb_medium_create.argtypes = [POINTER(StbMediumDesc), POINTER(c_void_p)]
    lib.stb_medium_create.restype = c_int
    lib.stb_medium_destroy.argtypes = [c_void_p]
    lib.stb_medium_destroy.restype = c_int
    lib.stb_medium_table.argtypes = [c_void_p, c_int, c_void_p, c_size_t]
    lib.stb_medium_table.restype = c_int
    lib.stb_trace_rk.argtypes = [c_void_p, POINTER(StbRk), c_void_p]
    lib.stb_trace_rk.restype = c_int
    for name in ("stb_device_count", "stb_set_device", "stb_field_create", "stb_field_destroy",
                 "stb_field_shape", "stb_field_peek", "stb_trace", "stb_check", "stb_trace_batch",
                 "stb_measure_fp64_peak", "stb_measure_store_bw"):
        getattr(lib, name).restype = c_int
    if lib.stb_abi_version() != STB_ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.stb_abi_version()} != {STB_ABI_VERSION}")
    return lib


lib = _load()
LIB_PATH = os.environ.get("SIDTRACE_B200_LIB", _build.LIB)


def check(rc: int) -> None:
    if rc != STB_OK:
        raise SidtraceError(rc, lib.stb_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib.stb_launch_count())
