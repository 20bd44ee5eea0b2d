"""Multi-GPU plumbing: one process per GPU, rays sharded by index, field replicated.

The path shards with NO data-path collective (rays are independent given the
read-only field).  Exchange steps exist only where the beam-level
caller has one (the reference's consumers -- Beam.field_r, beam.py:347-444;
field_result, beam.py:447-472 -- need every ray's O(1) summary in one place):
  * the sum over ranks of the flux-bin deposition histogram (extension),
  * the per-ray summaries (step counts, end states) of all shards in GLOBAL ray
    order on one rank (or all).
Full trajectories are never exchanged (10^8 rays x 10^3 steps x 32 B).

Transport.  On GPUs the exchange goes through the library's own multi-GPU entry
points (include/sidtrace_b200_comm.h: NCCL over NVLink 5 / NVSwitch, created from
a unique id; torch is the buffer owner only).  CPU tensors (the world-size-2
``gloo`` tests) go through torch.distributed.

Sharding.  ``mode="blocks"`` (default of ``trace_sharded``): the global index space
is cut into 256-ray blocks, block b belongs to rank b % world and is its local
block b // world.  Launch angles that reflect cluster in ray index, so every GPU
gets its share of them; a shard is a list of contiguous 256-ray runs, so putting
the gathered shards back in global order is one copy kernel over whole runs
(``stb_unshuffle_blocks``) -- round 1 interleaved single rays and paid an
``arange(n_global)`` + ``index_copy_`` per rank for it.  A shard is traced in
chunks; the gather of chunk k runs on a second stream while chunk k+1 is traced.
``mode="interleave"`` / ``"block"`` keep their round-1 meaning for ``shard_slice``.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

BLOCK = 256


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_slice(n_global: int, rank: int, world: int, mode: str = "interleave") -> slice:
    """Index slice of this rank's rays.  'interleave' deals single rays round-robin;
    'block' gives contiguous ranges.  (The block-interleaved sharding of ``trace_sharded`` is not
    a slice: see ``shard_index``.)"""
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
    if mode == "blocks":
        return len(shard_index(n_global, rank, world))
    return len(range(*shard_slice(n_global, rank, world, mode).indices(n_global)))


def shard_index(n_global: int, rank: int, world: int, block: int = BLOCK) -> np.ndarray:
    """Global ray indices of this rank's block-interleaved shard, in local order: blocks rank,
    rank + world, ... of `block` consecutive rays (the last global block may be partial)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    nb = -(-n_global // block)
    blocks = np.arange(rank, nb, world, dtype=np.int64)
    idx = (blocks[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n_global]


def local_blocks(n_global: int, world: int, block: int = BLOCK) -> int:
    """Local blocks of every rank's shard after padding: ceil(ceil(n/B) / world)."""
    nb = -(-n_global // block)
    return -(-nb // world)


# --------------------------------------------------------------------------
# transports
# --------------------------------------------------------------------------
class Comm:
    """The library's NCCL communicator (``stb_comm_*``).  ``Comm.from_torch()`` makes one for the
    current torch.distributed world: rank 0 draws the unique id, the process group only carries its
    128 bytes to the other ranks."""

    def __init__(self, unique_id: bytes, world: int, rank: int, device: int):
        from . import _cabi
        self._cabi = _cabi
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        self._h = ctypes.c_void_p()
        _cabi.check(_cabi.lib.stb_set_device(self.device))
        buf = ctypes.create_string_buffer(bytes(unique_id), _cabi.STB_COMM_ID_BYTES)
        _cabi.check(_cabi.lib.stb_comm_init(buf, self.world, self.rank, ctypes.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        from . import _cabi
        buf = ctypes.create_string_buffer(_cabi.STB_COMM_ID_BYTES)
        _cabi.check(_cabi.lib.stb_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, device=None) -> "Comm":
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.cuda.current_device() if device is None else torch.device(device).index
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], world, rank, dev)

    @staticmethod
    def _stream(stream=None):
        import torch
        s = torch.cuda.current_stream() if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    def allreduce_sum(self, t, stream=None):
        self._cabi.check(self._cabi.lib.stb_comm_allreduce_f64(self._h, t.data_ptr(), t.numel(), self._stream(stream)))
        return t

    def allreduce_max(self, t, stream=None):
        self._cabi.check(self._cabi.lib.stb_comm_allreduce_max_f64(self._h, t.data_ptr(), t.numel(), self._stream(stream)))
        return t

    def gather(self, send, recv, root=0, stream=None):
        """`send` (contiguous) from every rank into recv[world, ...] on `root`."""
        nbytes = send.numel() * send.element_size()
        self._cabi.check(self._cabi.lib.stb_comm_gather(self._h, root, send.data_ptr(), nbytes,
                                                        recv.data_ptr() if recv is not None else None, self._stream(stream)))

    def allgather(self, send, recv, stream=None):
        nbytes = send.numel() * send.element_size()
        self._cabi.check(self._cabi.lib.stb_comm_allgather(self._h, send.data_ptr(), nbytes, recv.data_ptr(), self._stream(stream)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._cabi.lib.stb_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _torch_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def allreduce_hist(hist, comm: Optional[Comm] = None):
    """In-place sum of the per-rank deposition histograms (one tiny latency-bound collective)."""
    if comm is not None and comm.world > 1 and hist.is_cuda:
        return comm.allreduce_sum(hist)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def gather_sharded(local, n_global: int, mode: str = "interleave", dim: int = -1):
    """All-gather (torch.distributed) a per-ray tensor whose `dim` axis is this rank's shard and put
    it back in global ray order: ONE unpadded-or-padded ``all_gather_into_tensor`` and a strided view
    per rank -- no index tensors.  Modes: 'interleave', 'block', 'blocks'."""
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
    src = local.movedim(dim, 0)                     # shard axis first
    rest = tuple(src.shape[1:])
    buf = torch.zeros((mx,) + rest, dtype=local.dtype, device=local.device)
    buf[:lens[rank]].copy_(src)
    parts = torch.empty((world, mx) + rest, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(parts.view((world * mx,) + rest), buf)
    out = torch.empty((n_global,) + rest, dtype=local.dtype, device=local.device)
    if mode == "blocks":
        for r in range(world):
            out[torch.from_numpy(shard_index(n_global, r, world)).to(out.device)] = parts[r, :lens[r]]
    else:
        for r in range(world):
            out[shard_slice(n_global, r, world, mode)] = parts[r, :lens[r]]      # strided / contiguous slice assignment
    return out.movedim(0, dim)


def stage_shard(launch, amp=None, *, mode="blocks"):
    """Copy this rank's shard of a global launch set (host numpy arrays, identical on every rank)
    into HBM once.  Returns (device launch tensors, device amp or None, n_global)."""
    import torch
    rank, world, local_rank = env_world()
    n_global = len(launch[0])
    sl = shard_index(n_global, rank, world) if mode == "blocks" else shard_slice(n_global, rank, world, mode)
    dev = torch.device("cuda", local_rank)
    d_in = [torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev) for a in launch]
    d_amp = torch.from_numpy(np.ascontiguousarray(amp[sl])).to(dev) if amp is not None else None
    return d_in, d_amp, n_global


@dataclass
class Gathered:
    """Per-ray summaries of the WHOLE beam in global ray order (on the root, or on every rank)."""
    nsteps: object = None        # int32 [n_global]
    end_state: object = None     # float64 [4, n_global]


@dataclass
class ShardTiming:
    """CUDA events around the pieces of one ``trace_sharded`` pass (device time, this rank)."""
    start: object = None
    traced: object = None        # last trace kernel done (main stream)
    done: object = None          # exchange done (main stream has joined the exchange stream)

    def ms(self):
        return {"kernel_ms": self.start.elapsed_time(self.traced), "total_ms": self.start.elapsed_time(self.done),
                "exchange_exposed_ms": self.traced.elapsed_time(self.done)}


def trace_sharded(field, launch, *, L, st, dx, dy, amp=None, nbins=0, hist_kw=None, gather=False,
                  mode="blocks", staged=None, chunks=4, comm: Optional[Comm] = None, root=0, timing=False, **kw):
    """Trace this rank's shard of a global launch set and run the exchange steps.  ``launch`` /
    ``amp`` are global host arrays (identical on every rank), or pass ``staged=stage_shard(...)``
    (same ``mode``) to reuse a shard already resident in HBM.

    gather: False; True / "root" -- step counts and end states of the whole beam, in global ray order,
    on rank ``root``; "all" -- on every rank.  With ``mode="blocks"`` and a ``Comm`` the shard is traced
    in ``chunks`` pieces and the exchange of a piece (NCCL send/recv on a second stream + the
    unshuffle kernel) overlaps the tracing of the next one.

    Returns (BatchResult of the local shard [nsteps always; end_state when not chunked], global hist or
    None, Gathered or None[, ShardTiming])."""
    import torch
    from . import _cabi, sidtrace
    d_in, d_amp, n_global = staged if staged is not None else stage_shard(launch, amp, mode=mode)
    world, rank = (comm.world, comm.rank) if comm is not None else _torch_world()
    dev = d_in[0].device
    n_loc = int(d_in[0].shape[0])
    hkw = dict(hist_kw or {})
    want_gather = bool(gather) and world > 1
    tm = ShardTiming() if timing else None
    main = torch.cuda.current_stream(dev)
    if tm:
        tm.start = torch.cuda.Event(enable_timing=True); tm.traced = torch.cuda.Event(enable_timing=True)
        tm.done = torch.cuda.Event(enable_timing=True)
        tm.start.record(main)

    chunked = want_gather and mode == "blocks" and comm is not None
    if not chunked:
        res = sidtrace.trace_batch(field, *d_in, L=L, st=st, dx=dx, dy=dy, amp=d_amp, nbins=nbins, **hkw, **kw)
        if tm:
            tm.traced.record(main)
        hist = allreduce_hist(res.hist, comm) if res.hist is not None else None
        gathered = None
        if gather and world == 1:
            gathered = Gathered(res.nsteps, res.end_state)
        elif want_gather:
            gathered = Gathered(gather_sharded(res.nsteps, n_global, mode), gather_sharded(res.end_state, n_global, mode, dim=1))
        if tm:
            tm.done.record(main)
        return (res, hist, gathered, tm) if timing else (res, hist, gathered)

    # ---- chunked trace with the exchange of chunk k under the trace of chunk k+1 ------------------
    B = BLOCK
    lb = local_blocks(n_global, world, B)                 # padded local blocks, same on every rank
    nchunks = max(1, min(int(chunks), lb))
    cb = -(-lb // nchunks)                                # local blocks per chunk
    nchunks = -(-lb // cb)
    cnp = cb * B                                          # padded rays per chunk
    to_all = gather == "all"
    i_am_root = to_all or rank == root
    hist = torch.zeros(nbins, dtype=torch.float64, device=dev) if nbins else None
    nsteps_loc = torch.zeros(nchunks * cnp, dtype=torch.int32, device=dev)
    end_loc = [torch.zeros((4, cnp), dtype=torch.float64, device=dev) for _ in range(nchunks)]
    g = Gathered()
    if i_am_root:
        g.nsteps = torch.empty(n_global, dtype=torch.int32, device=dev)
        g.end_state = torch.empty((4, n_global), dtype=torch.float64, device=dev)
        rbuf_n = torch.empty((world, cnp), dtype=torch.int32, device=dev)
        rbuf_e = torch.empty((world, 4, cnp), dtype=torch.float64, device=dev)
    xs = _exchange_stream(dev)
    lib = _cabi.lib
    # the chunks are traced on two alternating side streams: chunk c+1's CTAs fill the SMs while chunk c's last
    # wave drains, instead of every chunk paying its own tail in stream order (they write disjoint outputs and
    # add to the histogram with atomics)
    ts = _trace_streams(dev)
    ev0 = torch.cuda.Event()
    ev0.record(main)
    for t in ts:
        t.wait_event(ev0)
    for c in range(nchunks):
        lo = c * cnp
        hi = min(lo + cnp, n_loc)
        ns_c = nsteps_loc[lo:lo + cnp]
        if hi > lo:
            out = sidtrace.BatchResult(nsteps=ns_c[:hi - lo], end_state=None, hist=hist)
            # end states: the kernel wants [4][n] contiguous; a chunk that is not full writes a compact
            # [4][hi-lo] block which is spread to the padded rows below
            full = hi - lo == cnp
            e_c = end_loc[c] if full else torch.empty((4, hi - lo), dtype=torch.float64, device=dev)
            out.end_state = e_c
            with torch.cuda.stream(ts[c % 2]):
                sidtrace.trace_batch(field, *[a[lo:hi] for a in d_in], L=L, st=st, dx=dx, dy=dy,
                                     amp=d_amp[lo:hi] if d_amp is not None else None, hist=hist, out=out, **hkw, **kw)
                if not full:
                    end_loc[c][:, :hi - lo].copy_(e_c)
                    e_c.record_stream(ts[c % 2])          # allocated on the main stream, used on this one
        ev = torch.cuda.Event()
        ev.record(ts[c % 2])
        if c >= nchunks - 2:
            main.wait_event(ev)                           # main resumes after the last chunk of either stream
        if tm and c == nchunks - 1:
            tm.traced.record(main)
        xs.wait_event(ev)
        with torch.cuda.stream(xs):
            if to_all:
                comm.allgather(ns_c, rbuf_n, stream=xs)
                comm.allgather(end_loc[c], rbuf_e, stream=xs)
            else:
                comm.gather(ns_c, rbuf_n if i_am_root else None, root=root, stream=xs)
                comm.gather(end_loc[c], rbuf_e if i_am_root else None, root=root, stream=xs)
            if i_am_root:
                g0 = c * cb * world * B
                sp = ctypes.c_void_p(xs.cuda_stream)
                _cabi.check(lib.stb_unshuffle_blocks(rbuf_n.data_ptr(), g.nsteps.data_ptr(), world, 1, cb, B, 4, g0, n_global, n_global, sp))
                _cabi.check(lib.stb_unshuffle_blocks(rbuf_e.data_ptr(), g.end_state.data_ptr(), world, 4, cb, B, 8, g0, n_global, n_global, sp))
    if hist is not None:
        comm.allreduce_sum(hist, stream=main)
    fin = torch.cuda.Event()
    fin.record(xs)
    main.wait_event(fin)
    if tm:
        tm.done.record(main)
    res = sidtrace.BatchResult(nsteps=nsteps_loc[:n_loc], end_state=None, hist=hist)
    gathered = g if i_am_root else None
    return (res, hist, gathered, tm) if timing else (res, hist, gathered)


_xstreams: dict = {}
_tstreams: dict = {}


def _trace_streams(dev):
    """Two side streams per device for the chunk traces of ``trace_sharded``."""
    import torch
    key = (dev.type, dev.index)
    s = _tstreams.get(key)
    if s is None:
        s = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        _tstreams[key] = s
    return s


def _exchange_stream(dev):
    """One high-priority side stream per device for the exchange steps."""
    import torch
    key = (dev.type, dev.index)
    s = _xstreams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev, priority=-1)
        _xstreams[key] = s
    return s


def trace_rk_sharded(medium, launch, *, amp=None, xscale=None, yscale=None, gather=False, mode="interleave", **kw):
    """The adaptive-RK extension (physics.trace_rk) over this rank's shard: ``launch`` = global host
    arrays (x0, y0, px0, py0), identical on every rank; ``medium`` is this rank's replica.  The
    deposition histogram is all-reduced; with ``gather`` the per-ray (status, end_state) come back
    in global ray order.  Returns (RkResult of the local shard, global hist or None, gathered or None)."""
    from . import physics
    per_ray = [a for a in (amp, xscale, yscale)]
    d_in, _, n_global = stage_shard(launch, mode=mode)
    rank, world, _ = env_world()
    sl = shard_index(n_global, rank, world) if mode == "blocks" else shard_slice(n_global, rank, world, mode)
    import torch
    dev = d_in[0].device
    d_amp, d_xs, d_ys = (None if a is None else torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev) for a in per_ray)
    res = physics.trace_rk(medium, *d_in, amp=d_amp, xscale=d_xs, yscale=d_ys, **kw)
    hist = allreduce_hist(res.hist) if res.hist is not None else None
    gathered = None
    if gather:
        gathered = (gather_sharded(res.status, n_global, mode), gather_sharded(res.end_state, n_global, mode, dim=1))
    return res, hist, gathered
