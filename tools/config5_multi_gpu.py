#!/usr/bin/env python
"""BASELINE config 5 shape on N GPUs (torchrun): a big synthetic beam sharded by ray index
(256-ray blocks round-robin), endpoint + deposition-histogram output only, ONE NCCL all-reduce of
the flux-bin histogram and the per-ray summaries of the whole beam gathered in global ray order on
rank 0, the exchange of chunk k under the trace of chunk k+1 (rayrtrace_b200.dist.trace_sharded over
the library's stb_comm_* entry points).  --check compares the sharded result with a single-GPU run of
the whole beam on rank 0 (bit-identical step counts and end states; histogram within 1e-6 L1).
Prints device-timed throughput (max over ranks) with the kernel / exposed-exchange split.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config5_multi_gpu.py --nrays 8000000
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayrtrace_b200 import dist as rd, sidtrace, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nrays", type=int, default=8_000_000, help="global ray count")
    ap.add_argument("--nx", type=int, default=201)
    ap.add_argument("--nbins", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--check", action="store_true", help="rank 0 also traces the whole beam alone and compares")
    args = ap.parse_args()
    rank, world, local = rd.env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = rd.Comm.from_torch(dev)
    # every rank builds ONLY its block-interleaved shard of the fan, launch states prepared on the device
    idx = rd.shard_index(args.nrays, rank, world)
    case = synth.beam_case(nx=args.nx, nrays=args.nrays, seed=0, index=idx, device_launch=True)
    nf = case.nfield
    field = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny)
    hist_kw = dict(xorg=float(nf.x[0]), yorg=float(nf.y[0]), r_lcs=synth.R_DIA, rho_max=1.2)
    kw = dict(L=case.L, st=case.st, dx=case.dx, dy=case.dy)

    staged = ([torch.from_numpy(a).to(dev) for a in case.launch], torch.from_numpy(case.amp).to(dev), args.nrays)

    def run(gather="root"):
        return rd.trace_sharded(field, None, staged=staged, nbins=args.nbins, hist_kw=hist_kw, gather=gather, comm=comm,
                                chunks=args.chunks, refill=False, timing=True, **kw)

    for _ in range(2):
        res, hist, gathered, tm = run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    tms = []
    for _ in range(args.reps):
        res, hist, gathered, tm = run()
        tms.append(tm)
    e1.record()
    torch.cuda.synchronize()
    split = np.array([[t.ms()[k] for k in ("kernel_ms", "exchange_exposed_ms")] for t in tms]).mean(0)
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps, split[0], split[1]], dtype=torch.float64, device=dev)
    steps = res.nsteps.sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps, op=dist.ReduceOp.SUM)
    ok = None
    if args.check and rank == 0:
        whole = synth.beam_case(nx=args.nx, nrays=args.nrays, seed=0, index=np.arange(args.nrays, dtype=np.int64),
                                device_launch=True)
        d_in = [torch.from_numpy(a).to(dev) for a in whole.launch]
        one = sidtrace.trace_batch(field, *d_in, amp=torch.from_numpy(whole.amp).to(dev), nbins=args.nbins, **hist_kw, **kw)
        torch.cuda.synchronize()
        same_steps = bool(torch.equal(one.nsteps, gathered.nsteps))
        same_end = bool(torch.equal(one.end_state.view(torch.int64), gathered.end_state.view(torch.int64)))
        l1 = float((one.hist - hist).abs().sum() / one.hist.abs().sum())
        ok = {"nsteps_bit_identical": same_steps, "end_state_bit_identical": same_end, "hist_rel_l1": l1}
        assert same_steps and same_end and l1 < 1e-6, ok
    if rank == 0:
        print(json.dumps({"config": f"config5-shape: {args.nrays} rays, {args.nx}^2 grid, endpoints + {args.nbins}-bin deposition histogram",
                          "n_gpus": world, "ms_per_pass_incl_allreduce_and_gather": float(ms[0].item()),
                          "kernel_ms": float(ms[1].item()), "exchange_exposed_ms": float(ms[2].item()),
                          "ray_steps": float(steps.item()), "ray_steps_per_s": float(steps.item()) / (float(ms[0].item()) * 1e-3),
                          "hist_sum": float(hist.sum().item()), "check": ok}), flush=True)
    if world > 1:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
