"""Throughput of the adaptive-RK extension kernel (include/sidtrace_b200_rk.h) on one B200, with its
CPU checker timed beside it.  Not the headline bench (bench.py measures the reference's hot path);
this is the measurement of SURVEY 8 f-4's extension.

    python tools/rk_bench.py [--nrays 1000000] [--nx 1001] [--mode O|X] [--interp bspline] [--rtol 1e-8]

Prints one JSON line per configuration: step attempts/s (accepted + rejected Dormand-Prince steps),
right-hand-side evaluations/s (6 per attempt), accepted steps/s, and the CPU checker's rate.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def medium_tables(nx, absorb):
    xs = np.linspace(-10, 10, nx)
    XX, YY = np.meshgrid(xs, xs)
    r2 = (XX / 7.85) ** 2 + (YY / 7.85) ** 2
    X = 0.6 * np.exp(-2.5 * r2)                       # FT-2-sized Gaussian core, peak n = 0.6 n_c
    alpha = 0.05 * np.exp(-((np.sqrt(r2) - 0.5) / 0.1) ** 2) if absorb else None   # absorbing shell at rho = 0.5
    return xs, X, alpha


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nrays", type=int, default=1_000_000)
    ap.add_argument("--nx", type=int, default=1001)
    ap.add_argument("--mode", default="O", help="O, X (analytic 1/R field) or XT (X mode, |B| from a tabulated Solov'ev equilibrium)")
    ap.add_argument("--interp", default="bspline")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--absorb", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-rays", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lane-refill", action="store_true", help="lane-granular refill (refill_min = 1)")
    ap.add_argument("--refill-min", type=int, default=None, help="idle lanes that trigger a refill (32 = warp tiles; default 24)")
    ap.add_argument("--sort-rays", action="store_true", help="trace in Morton order of launch cell and direction")
    ap.add_argument("--shuffled", action="store_true", help="random launch order and angles: lanes of a warp touch unrelated cells")
    args = ap.parse_args()

    import torch
    from rayrtrace_b200 import physics as ph
    torch.cuda.set_device(0)
    xs, X, alpha = medium_tables(args.nx, args.absorb)
    eq = ph.Equilibrium.solovev(kappa=1.2, psi_b=6e-3, nr=129, nz=129, half=0.12) if args.mode == "XT" else None
    med = ph.Medium(X, xs, xs, alpha=alpha, mode="X" if args.mode == "XT" else args.mode, Y0=0.45, R0=55.0, interp=args.interp,
                    equilibrium=eq, freq=137.5e9 if eq else None)
    n = args.nrays
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = (torch.rand(n, generator=g, dtype=torch.float64, device="cuda") - 0.5) * 10
    if not args.shuffled:
        x0 = torch.sort(x0).values            # a beam: neighbouring ray indices are neighbouring rays
    y0 = torch.full_like(x0, -9.5)
    th = (torch.rand(n, generator=g, dtype=torch.float64, device="cuda") - 0.5) * (0.6 if args.shuffled else 0.02)
    if not args.shuffled:
        th = th + 0.03 * x0                   # a diverging fan with a little jitter
    N = ph.refractive_index(med, x0, y0)
    px, py = (N * torch.sin(th)).contiguous(), (N * torch.cos(th)).contiguous()
    kw = dict(rtol=args.rtol, atol_r=args.rtol, atol_p=args.rtol, hmax=0.4, lane_refill=args.lane_refill, refill_min=args.refill_min, sort_rays=args.sort_rays)
    if args.absorb:
        kw.update(nbins=64, r_lcs=7.85, rho_max=2.0)
    res = ph.RkResult()
    for _ in range(2):
        if res.hist is not None:
            res.hist.zero_()
        ph.trace_rk(med, x0, y0, px, py, out=res, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        ph.trace_rk(med, x0, y0, px, py, out=res, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    acc, rej = int(res.nacc.sum()), int(res.nrej.sum())
    attempts = acc + rej
    line = {
        "kernel": f"stbrk::rk_kernel<{ph.INTERP[args.interp]},{dict(O=0, X=1, XT=2)[args.mode]},{int(args.absorb)}>",
        "workload": f"{n} rays, {args.nx}^2 {args.interp} medium, {args.mode} mode, rtol {args.rtol:g}"
                    + (", absorption + 64-bin deposition" if args.absorb else "")
                    + (", shuffled launch order" if args.shuffled else ", rays ordered along the launch line")
                    + (", traced in Morton order" if args.sort_rays else "")
                    + (", lane refill" if args.lane_refill else ", refill_min %s" % (args.refill_min or "24 (default)")),
        "ms_per_launch": ms, "accepted": acc, "rejected": rej,
        "attempts_per_s": attempts / (ms * 1e-3), "accepted_steps_per_s": acc / (ms * 1e-3),
        "rhs_evals_per_s": (6 * attempts + n) / (ms * 1e-3),
        "max_abs_hamiltonian": float(res.ham.abs().max()),
        "status_counts": torch.bincount(res.status, minlength=5).tolist(),
    }
    if not args.no_cpu:
        from oracle import rk
        m = min(args.cpu_rays, n)
        y2 = None
        if eq is not None:
            y2 = rk.equilibrium_y2(dict(psi=eq.psi, fpol=eq.fpol, rleft=eq.rleft, rdim=eq.rdim, zmid=eq.zmid, zdim=eq.zdim,
                                        simagx=eq.simagx, sibdry=eq.sibdry, r0_cm=55.0, z0_cm=0.0,
                                        y_per_tesla=ph.ELECTRON_CYCLOTRON_HZ_PER_T / 137.5e9), args.nx, args.nx, med.xorg, med.yorg, med.dx, med.dy)
        om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, alpha=alpha, mode=args.mode, Y0=0.45, R0=55.0,
                       interp=ph.INTERP[args.interp], Y2=y2)
        h = [a[:m].cpu().numpy() for a in (x0, y0, px, py)]
        cores = os.cpu_count() or 1
        ckw = dict(kw, h0=0.5 * min(med.dx, med.dy), threads=cores)
        ckw.pop("lane_refill"); ckw.pop("refill_min"); ckw.pop("sort_rays")
        rk.trace(om, *[a[:256] for a in h], **ckw)
        t0 = time.perf_counter()
        w = rk.trace(om, *h, **ckw)
        dt = time.perf_counter() - t0
        same = w["end_state"].tobytes() == res.end_state[:, :m].cpu().numpy().tobytes()
        line["cpu_checker"] = {"attempts_per_s": float(w["nacc"].sum() + w["nrej"].sum()) / dt, "cores": cores,
                               "sample_rays": m, "bitwise_equal_to_gpu": bool(same)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
