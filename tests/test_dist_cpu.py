"""World-size-2 (and 3) gloo tests of the multi-GPU plumbing on CPU: the shard /
gather / all-reduce logic of rayrtrace_b200.dist, with the oracle standing in
for the kernels (this is a test: it may)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import random_field, random_launch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from rayrtrace_b200 import dist as rd
    rng = np.random.default_rng(5)
    nx, ny, n, L = 40, 30, (101 if mode != "blocks" else 1337), 80          # odd ray counts: ragged shards / a partial last block
    ax, ay = random_field(rng, nx, ny)
    launch = random_launch(rng, n, nx, ny)
    amp = rng.uniform(0.5, 1, n)
    sl = rd.shard_index(n, rank, world) if mode == "blocks" else rd.shard_slice(n, rank, world, mode)
    loc = oracle.trace_batch(*[a[sl] for a in launch], L, 0.3, 1.0, 1.0, nx, ny, ax, ay)
    hist = np.zeros(16)
    for r in range(len(loc["nsteps"])):
        oracle.deposit(loc["traj"][r, :loc["nsteps"][r]], 0.3, amp[sl][r], 0.0, 0.0, 1.0, 1.0, 20.0, 1.5, 16, hist)
    h = rd.allreduce_hist(torch.from_numpy(hist))
    ns = rd.gather_sharded(torch.from_numpy(loc["nsteps"]), n, mode)
    end = rd.gather_sharded(torch.from_numpy(loc["end"]), n, mode, dim=1)
    if rank == 0:
        q.put((h.numpy().copy(), ns.numpy().copy(), end.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "interleave"), (2, "block"), (3, "interleave"), (2, "blocks"), (3, "blocks")])
def test_sharded_equals_single(world, mode):
    import oracle
    oracle.port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    h, ns, end = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    nx, ny, n, L = 40, 30, (101 if mode != "blocks" else 1337), 80
    ax, ay = random_field(rng, nx, ny)
    launch = random_launch(rng, n, nx, ny)
    amp = rng.uniform(0.5, 1, n)
    one = oracle.trace_batch(*launch, L, 0.3, 1.0, 1.0, nx, ny, ax, ay)
    hist = np.zeros(16)
    for r in range(n):
        oracle.deposit(one["traj"][r, :one["nsteps"][r]], 0.3, amp[r], 0.0, 0.0, 1.0, 1.0, 20.0, 1.5, 16, hist)
    assert np.array_equal(ns, one["nsteps"])                 # rays are independent: bit-identical
    assert end.tobytes() == one["end"].tobytes()
    assert np.abs(h - hist).sum() <= 1e-12 * np.abs(hist).sum()   # summation order differs across shards


def test_shard_slices_partition():
    from rayrtrace_b200.dist import shard_len, shard_slice
    for n in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            for mode in ("interleave", "block"):
                seen = np.zeros(n, int)
                for r in range(world):
                    seen[shard_slice(n, r, world, mode)] += 1
                    assert shard_len(n, r, world, mode) == len(np.arange(n)[shard_slice(n, r, world, mode)])
                assert (seen == 1).all()
    with pytest.raises(ValueError):
        shard_slice(10, 2, 2)


def test_block_interleaved_sharding_matches_the_library():
    """dist.shard_index / local_blocks against the C entry points stb_shard_local_rays / _blocks, and the
    block-interleaved shards partition the rays."""
    from rayrtrace_b200 import _cabi
    from rayrtrace_b200.dist import local_blocks, shard_index
    lib = _cabi.lib
    for n in (0, 1, 255, 256, 257, 1000, 4096, 100_003):
        for world in (1, 2, 3, 8):
            seen = np.zeros(n, int)
            for r in range(world):
                idx = shard_index(n, r, world)
                seen[idx] += 1
                assert len(idx) == lib.stb_shard_local_rays(n, r, world, 256)
                assert len(idx) <= local_blocks(n, world) * 256
                if len(idx):
                    b = idx // 256
                    assert np.all(b % world == r) and np.all(np.diff(idx) > 0)
            assert (seen == 1).all()
            assert local_blocks(n, world) == lib.stb_shard_local_blocks(n, world, 256)


def test_unshuffle_formula_on_host():
    """The index map stb_unshuffle_blocks implements (local block lb of rank r -> global block lb*world + r),
    restated with numpy: gathering every rank's padded shard and unshuffling gives the identity."""
    from rayrtrace_b200.dist import local_blocks, shard_index
    B = 256
    for n, world in ((1000, 3), (5000, 2), (256 * 8 + 5, 8)):
        lb = local_blocks(n, world)
        src = np.full((world, lb * B), -1, np.int64)
        for r in range(world):
            idx = shard_index(n, r, world)
            src[r, :len(idx)] = idx                       # each ray carries its own global index
        dst = np.full(n, -2, np.int64)
        for r in range(world):
            for l in range(lb):
                g0 = (l * world + r) * B
                for j in range(B):
                    if g0 + j < n:
                        dst[g0 + j] = src[r, l * B + j]
        assert np.array_equal(dst, np.arange(n))
