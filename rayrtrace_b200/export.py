"""Trajectory export to host memory: trace a beam chunk by chunk and stream the KEPT part of every
trajectory (what the reference's Ray holds after ray.py:99-102) to the host in packed form,

    states[offsets[r] : offsets[r + 1]]  ==  the [nsteps[r], 4] array (x, y, kx, ky) of ray r,

with the tracing of chunk k+1 overlapped with the packing + device-to-host copy of chunk k on a
second stream.  This is the path for callers that really need every state on the host (the
reference keeps 3.2 MB of NumPy arrays per ray); it is PCIe-bound (~32 B per ray-step), which is
why everything else in this package keeps trajectories in HBM and moves O(rays) numbers only.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi, sidtrace


def trace_to_host(field, x0, y0, kx0, ky0, *, L, st, dx, dy, chunk_rays=131072, out=None, st_ray=None):
    """x0..ky0: host launch states (numpy, grid units).  Returns (offsets int64[n+1], states
    float64[total, 4] (pinned unless ``out`` is given), nsteps int32[n], end_state float64[4, n]).

    ``out``: optional preallocated [>= total, 4] float64 host array (pinned for full PCIe speed);
    when omitted a first endpoint-only pass sizes it exactly."""
    import torch
    n = int(x0.shape[0])
    dev = torch.device("cuda", torch.cuda.current_device())
    kw = dict(L=int(L), st=st, dx=dx, dy=dy)
    ins = [np.ascontiguousarray(a, np.float64) for a in (x0, y0, kx0, ky0)]
    if st_ray is not None:
        st_ray = np.ascontiguousarray(st_ray, np.float64)

    # pass 1 (cheap, endpoint mode): step counts -> offsets and the exact size of the export
    first = sidtrace.trace_batch(field, *ins, st_ray=st_ray, **kw)
    nsteps = np.asarray(first.nsteps)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(nsteps, out=offsets[1:])
    total = int(offsets[-1])
    if out is None:
        out = sidtrace.pinned_empty((max(total, 1), 4))
    if out.shape[0] < total or out.shape[1] != 4 or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float64 array of shape [>= {total}, 4]")

    chunk = int(min(max(chunk_rays, 128), max(n, 1)))
    chunk = -(-chunk // 256) * 256
    s_trace, s_copy = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    arenas = [torch.empty((chunk // 256, L, 256, 4), dtype=torch.float64, device=dev) for _ in range(2)]
    max_steps_chunk = int(max((offsets[min(i + chunk, n)] - offsets[i]) for i in range(0, n, chunk))) if n else 0
    packed = [torch.empty((max(max_steps_chunk, 1), 4), dtype=torch.float64, device=dev) for _ in range(2)]
    free = [torch.cuda.Event(), torch.cuda.Event()]        # arena/packed slot k is free again
    for e in free:
        e.record(s_copy)
    d_off = torch.from_numpy(offsets).to(dev)
    d_ns = torch.from_numpy(nsteps).to(dev)
    out_t = torch.from_numpy(out)
    torch.cuda.synchronize(dev)            # set-up copies were issued on the default stream
    for k, lo in enumerate(range(0, n, chunk)):
        hi = min(lo + chunk, n)
        m = hi - lo
        slot = k & 1
        with torch.cuda.stream(s_trace):
            d_in = [torch.from_numpy(a[lo:hi]).to(dev, non_blocking=True) for a in ins]
            d_sr = torch.from_numpy(st_ray[lo:hi]).to(dev, non_blocking=True) if st_ray is not None else None
            s_trace.wait_event(free[slot])
            traj = arenas[slot][: -(-m // 256)]
            sidtrace.trace_batch(field, *d_in, traj=traj, tiled=True, tile=256, st_ray=d_sr, want_end=False, stream=s_trace, **kw)
            done = torch.cuda.Event()
            done.record(s_trace)
        with torch.cuda.stream(s_copy):
            s_copy.wait_event(done)
            nst = int(offsets[hi] - offsets[lo])
            local_off = d_off[lo:hi] - int(offsets[lo])
            _cabi.check(_cabi.lib.stb_pack_traj(traj.data_ptr(), 256, 0, int(L), m, d_ns[lo:hi].data_ptr(),
                                                local_off.data_ptr(), packed[slot].data_ptr(),
                                                ctypes.c_void_p(s_copy.cuda_stream)))
            if nst:
                out_t[int(offsets[lo]): int(offsets[hi])].copy_(packed[slot][:nst], non_blocking=True)
            free[slot].record(s_copy)
            # keep the temporaries of this chunk alive until the copy stream is done with them
            traj.record_stream(s_copy); local_off.record_stream(s_copy)
    s_copy.synchronize()
    s_trace.synchronize()
    return offsets, out[:total], nsteps, np.asarray(first.end_state)
