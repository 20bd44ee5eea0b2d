"""Multi-GPU plumbing: one process per GPU, rays sharded by index, field replicated.

The path shards with NO data-path collective (rays are independent given the
read-only field).  Collectives appear only where the beam-level
caller has a real exchange step:
  * all-reduce (sum, FP64) of the flux-bin deposition histogram (extension),
  * all-gather of per-ray summaries (step counts, end states) when one rank
    needs the whole beam.
Full trajectories are never gathered (10^8 rays x 10^3 steps x 32 B).

torch.distributed is the transport (NCCL over NVLink on GPUs, gloo in the CPU
tests); tensors must live where the backend expects them.
"""
from __future__ import annotations

import os

import numpy as np


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_slice(n_global: int, rank: int, world: int, mode: str = "interleave") -> slice:
    """Index slice of this rank's rays.  'interleave' (default) deals rays round-robin so
    that reflected / trapped rays, which cluster in launch angle, spread evenly over GPUs;
    'block' gives contiguous ranges."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    if mode == "interleave":
        return slice(rank, n_global, world)
    if mode == "block":
        lo = rank * n_global // world
        hi = (rank + 1) * n_global // world
        return slice(lo, hi)
    raise ValueError(f"unknown shard mode {mode!r}")


def shard_len(n_global: int, rank: int, world: int, mode: str = "interleave") -> int:
    return len(range(*shard_slice(n_global, rank, world, mode).indices(n_global)))


def allreduce_hist(hist):
    """In-place sum of the per-rank deposition histograms (one tiny latency-bound collective)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def gather_sharded(local, n_global: int, mode: str = "interleave", dim: int = -1):
    """All-gather a per-ray tensor whose `dim` axis is this rank's shard and put it back in
    global ray order.  Shards may differ in length by one; they are padded for the collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    dim = dim % local.ndim
    lens = [shard_len(n_global, r, world, mode) for r in range(world)]
    if local.shape[dim] != lens[rank]:
        raise ValueError(f"rank {rank}: shard length {local.shape[dim]} != expected {lens[rank]}")
    mx = max(lens)
    pad_shape = list(local.shape)
    pad_shape[dim] = mx
    buf = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    buf.narrow(dim, 0, lens[rank]).copy_(local)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous())
    out_shape = list(local.shape)
    out_shape[dim] = n_global
    out = torch.empty(out_shape, dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = torch.arange(n_global, device=local.device)[shard_slice(n_global, r, world, mode)]
        out.index_copy_(dim, idx, parts[r].narrow(dim, 0, lens[r]))
    return out


def stage_shard(launch, amp=None, *, mode="interleave"):
    """Copy this rank's shard of a global launch set (host numpy arrays, identical on every rank)
    into HBM once.  Returns (device launch tensors, device amp or None, n_global)."""
    import torch
    rank, world, local_rank = env_world()
    n_global = len(launch[0])
    sl = shard_slice(n_global, rank, world, mode)
    dev = torch.device("cuda", local_rank)
    d_in = [torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev) for a in launch]
    d_amp = torch.from_numpy(np.ascontiguousarray(amp[sl])).to(dev) if amp is not None else None
    return d_in, d_amp, n_global


def trace_sharded(field, launch, *, L, st, dx, dy, amp=None, nbins=0, hist_kw=None, gather=False,
                  mode="interleave", staged=None, **kw):
    """Trace this rank's shard of a global launch set and run the exchange steps.  ``launch`` /
    ``amp`` are global host arrays (identical on every rank), or pass ``staged=stage_shard(...)``
    to reuse a shard already resident in HBM.  Returns (BatchResult of the local shard, global
    hist or None, gathered (nsteps, end_state) in global ray order or None)."""
    from . import sidtrace
    d_in, d_amp, n_global = staged if staged is not None else stage_shard(launch, amp, mode=mode)
    res = sidtrace.trace_batch(field, *d_in, L=L, st=st, dx=dx, dy=dy, amp=d_amp, nbins=nbins, **(hist_kw or {}), **kw)
    hist = allreduce_hist(res.hist) if res.hist is not None else None
    gathered = None
    if gather:
        gathered = (gather_sharded(res.nsteps, n_global, mode), gather_sharded(res.end_state, n_global, mode, dim=1))
    return res, hist, gathered


def trace_rk_sharded(medium, launch, *, amp=None, xscale=None, yscale=None, gather=False, mode="interleave", **kw):
    """The adaptive-RK extension (physics.trace_rk) over this rank's shard: ``launch`` = global host
    arrays (x0, y0, px0, py0), identical on every rank; ``medium`` is this rank's replica.  The
    deposition histogram is all-reduced; with ``gather`` the per-ray (status, end_state) come back
    in global ray order.  Returns (RkResult of the local shard, global hist or None, gathered or None)."""
    from . import physics
    per_ray = [a for a in (amp, xscale, yscale)]
    d_in, _, n_global = stage_shard(launch, mode=mode)
    rank, world, _ = env_world()
    sl = shard_slice(n_global, rank, world, mode)
    import torch
    dev = d_in[0].device
    d_amp, d_xs, d_ys = (None if a is None else torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev) for a in per_ray)
    res = physics.trace_rk(medium, *d_in, amp=d_amp, xscale=d_xs, yscale=d_ys, **kw)
    hist = allreduce_hist(res.hist) if res.hist is not None else None
    gathered = None
    if gather:
        gathered = (gather_sharded(res.status, n_global, mode), gather_sharded(res.end_state, n_global, mode, dim=1))
    return res, hist, gathered
