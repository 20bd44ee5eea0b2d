#!/usr/bin/env python
"""Times the trace kernel variants on one GPU (CUDA events, device-resident inputs).

    python tools/kernel_sweep.py [--nrays 1000000] [--nx 1001] [--reps 5]

One line per variant: layout x schedule x division x output mode, with ray-steps/s and the
HBM-roofline fraction (32 B per ray-step over MEASURED_PEAKS.json's copy bandwidth)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayrtrace_b200 import sidtrace, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nrays", type=int, default=1_000_000)
    ap.add_argument("--nx", type=int, default=1001)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--case", default="beam", choices=["beam", "sweep", "trapped"])
    ap.add_argument("--only", default="", help="comma-separated substrings; a variant runs if any matches")
    ap.add_argument("--hoist", default="on", choices=["on", "off", "both"], help="hoisted per-cell products (default) or plain cells")
    ap.add_argument("--cells", default="x", choices=["x", "y", "both"], help="cell-table layout: the reference's x-major rows, the y-major copy, or both")
    ap.add_argument("--sorted", action="store_true", help="order the rays by launch cell (Morton) and direction first (not timed)")
    ap.add_argument("--sort-rays", choices=["off", "timed", "reuse"], default="off",
                    help="trace_batch(sort_rays=True): the coherent order computed inside the timed region, or once and reused")
    ap.add_argument("--L", type=int, default=0, help="states per ray (0 = the case's default cap)")
    ap.add_argument("--store-pattern", action="store_true", help="only run the store-pattern microbenchmark")
    args = ap.parse_args()
    peak = 6533.5
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    case = {"beam": synth.beam_case, "sweep": synth.sweep_case, "trapped": synth.trapped_case}[args.case](nx=args.nx, nrays=args.nrays)
    dev = torch.device("cuda", 0)
    nf = case.nfield
    field = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny)
    if args.sorted:
        xg, yg, kx, ky = case.launch
        ix, iy = np.clip(xg.astype(np.int64), 0, 65535), np.clip(yg.astype(np.int64), 0, 65535)

        def spread(v):
            v = (v | (v << 8)) & 0x00FF00FF
            v = (v | (v << 4)) & 0x0F0F0F0F
            v = (v | (v << 2)) & 0x33333333
            return (v | (v << 1)) & 0x55555555
        ang = np.floor((np.arctan2(kx, ky) + np.pi) / (2 * np.pi) * 64).astype(np.int64)
        key = ((spread(ix >> 2) | (spread(iy >> 2) << 1)) << 6) | ang          # 4x4-cell blocks, then direction
        order = np.argsort(key, kind="stable")
        case.launch = tuple(np.ascontiguousarray(a[order]) for a in case.launch)
        case.amp, case.x0, case.y0 = (np.ascontiguousarray(a[order]) for a in (case.amp, case.x0, case.y0))
    d_in = [torch.from_numpy(a).to(dev) for a in case.launch]
    d_amp = torch.from_numpy(np.ascontiguousarray(case.amp)).to(dev)
    d_xp = torch.from_numpy(np.ascontiguousarray(case.x0)).to(dev)
    d_yp = torch.from_numpy(np.ascontiguousarray(case.y0)).to(dev)
    n, L = case.nrays, (args.L or case.L)
    if args.store_pattern:
        import ctypes
        lib = sidtrace._cabi.lib
        g = ctypes.c_double()
        sidtrace._cabi.check(lib.stb_measure_store_bw(8 << 30, ctypes.byref(g)))
        print(f"# plain grid-stride streaming stores: {g.value:.0f} GB/s")
        nt = (n + 255) // 256
        for tile, ntl in ((256, nt), (128, 2 * nt), (32, 8 * nt)):
            for steps in (820, 1000):
                for chain in (0, 4, 8, 16, 32):
                    sidtrace._cabi.check(lib.stb_measure_store_pattern(ntl, tile, steps, steps, chain, ctypes.byref(g)))
                    print(f"store-pattern tile={tile:3d} ntiles={ntl:6d} rows/tile={steps} chain={chain:2d}: {g.value:8.1f} GB/s  "
                          f"frac {g.value / peak:5.3f}", flush=True)
        # rows allocated at the cap (the current bench layout: L rows per tile, ~820 written)
        for chain in (0, 8):
            sidtrace._cabi.check(lib.stb_measure_store_pattern(nt, 256, L, 820, chain, ctypes.byref(g)))
            print(f"store-pattern tile=256 ntiles={nt} rows/tile={L} (820 written) chain={chain}: {g.value:8.1f} GB/s  frac {g.value / peak:5.3f}", flush=True)
        return
    nt = (n + 255) // 256 * 8
    buf = None   # trajectory arena, allocated when a trajectory variant is selected
    kw = dict(L=L, st=case.st, dx=case.dx, dy=case.dy)
    out = sidtrace.BatchResult(nsteps=torch.empty(n, dtype=torch.int32, device=dev),
                               end_state=torch.empty((4, n), dtype=torch.float64, device=dev))
    variants = []
    hoists = {"on": (True,), "off": (False,), "both": (True, False)}[args.hoist]
    for mode in ("traj-tiled", "traj-tiled64", "traj-tiled128", "traj-tiled256", "traj-ray", "traj-exact", "traj-pass2", "endpoints",
                 "endpoints-tiled", "extras", "extras-tiled", "hist64", "hist1024", "cross", "lookup"):
        for refill in (True, False):
            for div_ieee in (False, True):
                for hoist in hoists:
                    for cells in {"x": ("x",), "y": ("y",), "both": ("x", "y")}[args.cells]:
                        variants.append((mode, refill, div_ieee, hoist, cells))
    print(f"# case {case.name} L={L} peak={peak} GB/s lib={os.path.basename(sidtrace._cabi.LIB_PATH)}")
    first = True
    plan = None
    for mode, refill, div_ieee, hoist, cells in variants:
        tag = f"{mode}/{'refill' if refill else 'static'}/{'ieee' if div_ieee else 'divc'}/{'hoist' if hoist else 'cells'}" + ("/ymajor" if cells == "y" else "")
        if args.only and not any(o in tag for o in args.only.split(",")):
            continue
        o = sidtrace.BatchResult(nsteps=out.nsteps, end_state=out.end_state)
        k = dict(kw, refill=refill, div_ieee=div_ieee, hoist=hoist, cells=cells)
        if args.sort_rays == "timed":
            k["sort_rays"] = True
        elif args.sort_rays == "reuse":
            k["order"] = sidtrace.coherent_order(*d_in)
        if mode in ("traj-exact", "traj-pass2"):
            # two-pass exactly-sized arena: the whole of trace_exact, or its second pass alone (plan given)
            if refill:
                continue           # pass 1 picks lane refill itself, pass 2 is static
            del k["refill"]
            if plan is None:
                r0 = sidtrace.trace_batch(field, *d_in, out=sidtrace.BatchResult(nsteps=out.nsteps, end_state=out.end_state), **kw)
                plan = sidtrace.plan_trajectories(r0.nsteps, tile=256, L=L)
                arena = torch.empty((plan.total_rows, 256, 4), dtype=torch.float64, device=dev)
                print(f"# exact arena: {plan.arena_bytes / 1e9:.2f} GB for {int(r0.nsteps.sum())} ray-steps "
                      f"({plan.arena_bytes / 32 / max(int(r0.nsteps.sum()), 1):.4f} states allocated per ray-step)")
            if mode == "traj-pass2":
                fn = lambda: sidtrace.trace_batch(field, *d_in, traj=arena, plan=plan, want_end=False, out=o, **k)
            else:
                def fn():
                    r1 = sidtrace.trace_batch(field, *d_in, out=o, **k)
                    pl = sidtrace.plan_trajectories(r1.nsteps, tile=256, L=L)
                    assert pl.total_rows == plan.total_rows
                    sidtrace.trace_batch(field, *d_in, traj=arena, plan=pl, want_end=False,
                                         out=sidtrace.BatchResult(nsteps=o.nsteps), **k)
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            steps = int(o.nsteps.sum().item())
            ms = float(np.median(ts)); rate = steps / (ms * 1e-3)
            print(f"{tag:40s} {ms:9.3f} ms  {rate:11.4e} ray-steps/s  {rate * 32 / 1e9:8.1f} GB/s-equiv  frac {rate * 32 / 1e9 / peak:5.3f}  (min {min(ts):.3f} ms)", flush=True)
            continue
        if mode.startswith("traj") and buf is None:
            buf = torch.empty(nt * 32 * L * 4, dtype=torch.float64, device=dev)
        if mode in ("traj-tiled", "traj-tiled64", "traj-tiled128", "traj-tiled256"):
            T = {"traj-tiled": 32, "traj-tiled64": 64, "traj-tiled128": 128, "traj-tiled256": 256}[mode]
            ntt = (n + T - 1) // T
            k.update(traj=buf[: ntt * L * T * 4].view(ntt, L, T, 4), tiled=True, tile=T)
        elif mode == "traj-ray":
            k.update(traj=buf[: n * L * 4].view(n, L, 4))
        elif mode == "extras":
            k.update(want_extras=True)
        elif mode in ("hist64", "hist1024"):
            nb = int(mode[4:])
            k.update(hist=torch.zeros(nb, dtype=torch.float64, device=dev), amp=d_amp, xorg=float(nf.x[0]), yorg=float(nf.y[0]),
                     r_lcs=synth.R_DIA, rho_max=1.3)
        elif mode == "cross":
            k.update(cross=(9.0, "y"), x0_phys=d_xp, y0_phys=d_yp)
        elif mode == "lookup":
            c0 = sidtrace.trace_batch(field, *d_in, cross=(9.0, "y"), x0_phys=d_xp, y0_phys=d_yp, **kw)
            ph = c0.cross_state[5]
            k.update(cross=(9.0, "y"), x0_phys=d_xp, y0_phys=d_yp,
                     lookup_ph=torch.stack([torch.roll(ph, 1), torch.roll(ph, -1)]).contiguous())
        elif mode == "endpoints-tiled":
            k.update(tiled=True)
        elif mode == "extras-tiled":
            k.update(want_extras=True, tiled=True)
        for _ in range(2):
            sidtrace.trace_batch(field, *d_in, out=o, **k)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            sidtrace.trace_batch(field, *d_in, out=o, **k)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        steps = int(o.nsteps.sum().item())
        if first:
            ns = o.nsteps.cpu().numpy()
            print(f"# steps/ray: median {int(np.median(ns))} mean {ns.mean():.0f} p99 {int(np.percentile(ns, 99))} max {ns.max()} capped {(ns == L - 1).mean():.4f}")
            first = False
        ms = float(np.median(ts))
        rate = steps / (ms * 1e-3)
        frac = rate * 32 / 1e9 / peak
        print(f"{tag:40s} {ms:9.3f} ms  {rate:11.4e} ray-steps/s  {rate * 32 / 1e9:8.1f} GB/s-equiv  frac {frac:5.3f}  (min {min(ts):.3f} ms)", flush=True)


if __name__ == "__main__":
    main()
