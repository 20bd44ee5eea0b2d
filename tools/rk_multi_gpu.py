"""Adaptive-RK extension sharded over N GPUs (torchrun): checks the sharded result against one GPU
tracing the whole beam (bit-identical per-ray outputs, histogram to 1e-12) and reports attempts/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/rk_multi_gpu.py --nrays-per-gpu 1000000
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nrays-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--nx", type=int, default=1001)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from rayrtrace_b200 import dist as sdist, physics as ph
    from rk_bench import medium_tables
    rank, world, local = sdist.env_world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xs, X, alpha = medium_tables(args.nx, True)
    med = ph.Medium(X, xs, xs, alpha=alpha)
    n = args.nrays_per_gpu * world
    x0 = np.linspace(-5, 5, n)
    th = 0.03 * x0
    launch = (x0, np.full(n, -9.5), np.sin(th), np.cos(th))
    amp = np.exp(-(x0 / 3.0) ** 2)
    kw = dict(rtol=1e-8, atol_r=1e-8, atol_p=1e-8, hmax=0.4, nbins=64, r_lcs=7.85, rho_max=2.0)
    res, hist, gathered = sdist.trace_rk_sharded(med, launch, amp=amp, gather=True, **kw)
    torch.cuda.synchronize()
    # timing: per-rank shard resident, max over ranks
    sl = sdist.shard_slice(n, rank, world)
    d = [torch.from_numpy(np.ascontiguousarray(a[sl])).cuda() for a in launch]
    d_amp = torch.from_numpy(np.ascontiguousarray(amp[sl])).cuda()
    out = ph.RkResult()
    for _ in range(2):
        ph.trace_rk(med, *d, amp=d_amp, out=out, **kw)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out.hist.zero_()
        ph.trace_rk(med, *d, amp=d_amp, out=out, **kw)
        sdist.allreduce_hist(out.hist)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device="cuda")
    att = (out.nacc.sum() + out.nrej.sum()).to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(att, op=dist.ReduceOp.SUM)
    line = None
    if rank == 0:
        line = {"n_gpus": world, "nrays": n, "ms_per_pass": float(ms), "attempts_per_s": float(att) / (float(ms) * 1e-3)}
        if n <= 4_000_000:   # whole beam on one GPU as the check
            full = [torch.from_numpy(a).cuda() for a in launch]
            one = ph.trace_rk(med, *full, amp=torch.from_numpy(amp).cuda(), **kw)
            torch.cuda.synchronize()
            line["status_equal"] = bool(torch.equal(gathered[0], one.status))
            line["end_state_bitwise_equal"] = bool(torch.equal(gathered[1], one.end_state))
            line["hist_rel_l1"] = float((hist - one.hist).abs().sum() / one.hist.abs().sum())
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
