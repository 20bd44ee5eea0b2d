#!/usr/bin/env python
"""bench.py -- ray-steps/s (FP64) of the sidtrace hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic input: the
BASELINE config-3 workload (10^6-ray seven-beam FT-2 fan through a 1001^2
shaped-density grid) PER GPU, full-trajectory output ([ray][step][4] FP64 in
HBM) + per-ray step counts and end states.  Weak scaling: every rank traces
its own 10^6-ray interleaved shard of an N*10^6-ray fan; no data-path
collective (rays are independent).

The same line carries, measured in the same invocation:
  config5   BASELINE config 5 -- an N x 1.25*10^7-ray beam (10^8 rays at N = 8) on a 201^2 grid sharded
            over the N ranks, endpoints + 64-bin deposition histogram, NCCL all-reduce of the histogram and
            the per-ray summaries of the whole beam gathered in global ray order on rank 0
            (rayrtrace_b200.dist.trace_sharded): the sharded path WITH its exchange, max over ranks;
  endpoint_mode   the O(rays)-output mode of config 3 against the measured FP64 peak;
  exact_arena     config 3 through the two-pass exactly-sized trajectory arena (trace_exact);
  app             the reference's top-level driver shape (ft2.phases, ft2.py:25-71) end to end, wall clock:
                  3 density scales x 7 antenna pairs x 251 rays, and one 7 x 10^6-ray antenna scan;
  parity_check    a >= 10^4-ray subsample of the timed kernel's trajectories, bit for bit against the oracle
                  (untimed; the oracle is the checker, never the thing measured).

Prints ONE JSON line on rank 0 (contract: task statement / DESIGN.md section 6).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ray-steps/sec (FP64)"
UNIT = "ray-steps/s"
BYTES_PER_RAY_STEP = 32          # x, y, kx, ky FP64 written per step
FLOPS_PER_RAY_STEP = 16          # sidtracemodule.c:41-45


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nrays", type=int, default=1_000_000, help="rays per GPU")
    ap.add_argument("--nx", type=int, default=1001, help="grid nodes per axis")
    ap.add_argument("--cap", type=int, default=0, help="states per ray (0 = synth.default_cap)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget")
    ap.add_argument("--extra", action="store_true", help="also time endpoint-only mode and the FP64 peak")
    ap.add_argument("--no-traj-export", action="store_true", help="skip the trajectories-to-host side measurement")
    ap.add_argument("--no-config5", action="store_true", help="skip the sharded config-5 block")
    ap.add_argument("--config5-rays", type=int, default=12_500_000, help="config-5 rays per GPU (10^8 / 8)")
    ap.add_argument("--config5-chunks", type=int, default=8, help="config-5: pieces a shard is traced in (the exchange of a piece "
                    "hides under the trace of the next; only the last one is exposed)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed trajectory digest against the oracle")
    ap.add_argument("--no-app", action="store_true", help="skip the application-level (ft2.phases) block")
    ap.add_argument("--layout", default="tiled", choices=["tiled", "ray"],
                    help="trajectory layout: tiled [tile][step][32][4] (default) or ray-major [ray][step][4]")
    return ap.parse_args()


# --------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference extension (oracle/_ref), called per ray exactly as
# ray.py:80-83 does, on all host cores (the extension holds the GIL -> processes).
# --------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(nx, ny, ax, ay, L, kind):
    import oracle
    _W["fn"] = oracle.ref().trace if kind == "reference" else oracle.trace
    _W["ax"], _W["ay"], _W["nx"], _W["ny"] = ax, ay, nx, ny
    # prefaulted per-worker buffers: times the C loop, not np.zeros
    _W["buf"] = [np.zeros(L) for _ in range(4)]


def _cpu_worker_run(job):
    xg, yg, kx, ky, st, dx, dy = job
    x, y, bkx, bky = _W["buf"]
    fn, ax, ay, nx, ny = _W["fn"], _W["ax"], _W["ay"], _W["nx"], _W["ny"]
    total = 0
    for r in range(len(xg)):
        x[0] = xg[r]; y[0] = yg[r]; bkx[0] = kx[r]; bky[0] = ky[r]
        total += fn(x, y, bkx, bky, st, dx, dy, nx, ny, ax, ay)
    return total


class CpuArm:
    """Times the reference on a bounded sample of the workload."""

    def __init__(self, case, cores=None):
        import multiprocessing as mp
        import oracle
        self.kind = "reference" if oracle.have_ref() else "port"
        if self.kind == "port":
            oracle.port()
        self.cores = cores or len(os.sched_getaffinity(0))
        self.case = case
        nf = case.nfield
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init,
                             initargs=(nf.nx, nf.ny, np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), case.L, self.kind))

    def step(self, rays_per_core):
        """One bounded sample: rays_per_core * cores rays drawn (fixed seed) from the workload."""
        c = self.case
        n = rays_per_core * self.cores
        idx = np.random.default_rng(1234).choice(c.nrays, size=min(n, c.nrays), replace=False)
        parts = np.array_split(idx, self.cores)
        jobs = [(c.launch[0][p], c.launch[1][p], c.launch[2][p], c.launch[3][p], c.st, c.dx, c.dy) for p in parts]
        t0 = time.perf_counter()
        totals = self.pool.map(_cpu_worker_run, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        return int(sum(totals)), dt, len(idx)

    def close(self):
        self.pool.close()
        self.pool.join()


def calibrate_cpu(arm, budget_s):
    """Pick rays/core so one sample takes about budget_s."""
    steps, dt, n = arm.step(64)
    rate = steps / dt
    per_ray = max(steps / max(n, 1), 1.0)
    rays = int(max(64, min(rate * budget_s / per_ray / arm.cores, arm.case.nrays // arm.cores)))
    return rays


# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any: evidence from
    a profiling run of this command line, NOT a property of this run -- it is tagged with its source."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def workload_name(args):
    return (f"config3: {args.nrays}-ray seven-beam FT-2 fan per GPU, {args.nx}x{args.nx} shaped-density grid, "
            f"full trajectories f64 in HBM, layout {'[ray/256][step][ray%256][4]' if args.layout == 'tiled' else '[ray][step][4]'}")


def make_config(args, cap_states):
    """The `config` object of the JSON line -- byte-identical in both arms (--impl b200 / reference)."""
    return {"workload": workload_name(args), "nx": args.nx, "nrays_per_gpu": args.nrays, "cap_states": int(cap_states),
            "l2": "no explicit flush: every step streams ~26 GB of trajectory stores per GPU (>> 126 MB L2)",
            "parallelism": f"rays interleaved over {args.gpus} rank(s), field replicated, no data-path collective"}


# --------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from rayrtrace_b200 import synth
    case = synth.beam_case(nx=args.nx, nrays=args.nrays, seed=0, L=args.cap or None)
    arm = CpuArm(case)
    total_iters = max(args.steps + args.warmup, 1)
    per_step_budget = min(8.0, max(0.5, 150.0 / total_iters))
    rays = calibrate_cpu(arm, per_step_budget)
    for _ in range(args.warmup):
        arm.step(rays)
    tot_steps, tot_t, n = 0, 0.0, 0
    for _ in range(args.steps):
        s, dt, n = arm.step(rays)
        tot_steps += s
        tot_t += dt
    arm.close()
    value = tot_steps / tot_t
    sample = f"{n} rays/step drawn (seed 1234) from the {case.nrays}-ray workload, ~{tot_steps // max(args.steps, 1)} ray-steps/step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, case.L),
        "note": "reference C extension (sidtracemodule.c, gcc -O2) called per ray as ray.py:80-83 does, "
                "one process per host core, prefaulted MaxStep buffers",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def parity_digest(case, traj, nsteps_dev, tiled, sample=10_240):
    """Untimed: a >= 10^4-ray subsample of the trajectories the timed kernel just wrote, state for state
    against the oracle (the CPU restatement of sidtracemodule.c:36-48).  The oracle is the checker here,
    never the thing measured."""
    import hashlib
    import oracle
    import torch
    n = case.nrays
    m = min(n, sample)
    sel = np.unique(np.linspace(0, n - 1, m).astype(np.int64))
    nf = case.nfield
    ns_gpu = nsteps_dev.cpu().numpy()[sel]
    bad = 0
    h_gpu, h_cpu = hashlib.sha256(), hashlib.sha256()
    # chunks of 512 rays keep the host copy of the oracle's [rays, L, 4] block small
    for lo in range(0, len(sel), 512):
        part = sel[lo:lo + 512]
        w = oracle.trace_batch(*[np.ascontiguousarray(a[part]) for a in case.launch], case.L, case.st, case.dx, case.dy,
                               nf.nx, nf.ny, np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nthreads=4)
        if not np.array_equal(w["nsteps"], ns_gpu[lo:lo + 512]):
            bad += 1
            continue
        idx = torch.from_numpy(part).to(traj.device)
        if tiled:
            g = traj[idx // 256, :, idx % 256, :]            # [rays, L, 4]
        else:
            g = traj[idx]
        g = g.cpu().numpy()
        for r in range(len(part)):
            k = int(w["nsteps"][r]) + 1
            h_gpu.update(g[r, :k].tobytes())
            h_cpu.update(w["traj"][r, :k].tobytes())
    ok = bad == 0 and h_gpu.hexdigest() == h_cpu.hexdigest()
    return {"result": "ok" if ok else "MISMATCH", "rays": int(len(sel)), "states": int((ns_gpu.astype(np.int64) + 1).sum()),
            "sha256": h_gpu.hexdigest()[:16], "against": "oracle.trace_batch (CPU restatement of sidtracemodule.c:36-48), untimed"}


def run_app(_cabi):
    """ft2.phases (ft2.py:25-71) end to end on one GPU: Nfield per density scale, seven TX beams (wavefront,
    launch conditions, launch states), trace with fused receiver crossing, equal-phase lookup launch, ray-tube
    field, overlap with the RX pattern, Gaussian fit, interferometer phases.  Wall clock of the whole call."""
    import torch
    from rayrtrace_b200 import ft2, synth
    n, x, y = synth.shaped_density(201, shift=0.0, elon=1.0)
    ft2.phases(n, x, y, a_set=[0.5], nray=251)                      # warm-up: allocator, library, SciPy imports
    out = {}
    for key, a_set, nray, reps in (("phases_3x7x251", np.linspace(0.2, 1.0, 3), 251, 3), ("scan_7x1e6", [1.0], 1_000_003, 2)):
        best, launches = None, 0
        for _ in range(reps):
            torch.cuda.synchronize()
            l0 = _cabi.launch_count()
            t0 = time.perf_counter()
            ph_max, ph_sig = ft2.phases(n, x, y, a_set=a_set, nray=nray)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            launches = _cabi.launch_count() - l0
            best = dt if best is None else min(best, dt)
        out[key] = {"wall_s": best, "beams": 7 * len(a_set), "rays_per_beam": nray, "library_launches": int(launches),
                    "ph_sig_first": [round(float(v), 6) for v in ph_sig[-1, :3]]}
    out["note"] = ("rayrtrace_b200.ft2.phases on a 201^2 FT-2 profile, best wall time of a few calls; every O(rays) "
                   "array stays on the GPU (receiver_field_device / field_result_device)")
    return out


def run_config5(args, rank, world, local_rank, dev, comm):
    """BASELINE config 5 under the bench's clock: the ray-sharded beam WITH its exchange."""
    import torch
    import torch.distributed as dist
    from rayrtrace_b200 import dist as rd, sidtrace, synth
    n_global = args.config5_rays * world
    idx = rd.shard_index(n_global, rank, world)
    case = synth.beam_case(nx=201, nrays=n_global, seed=0, index=idx, device_launch=True)
    nf = case.nfield
    field = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny)
    hist_kw = dict(xorg=float(nf.x[0]), yorg=float(nf.y[0]), r_lcs=synth.R_DIA, rho_max=1.2)
    kw = dict(L=case.L, st=case.st, dx=case.dx, dy=case.dy)
    staged = ([torch.from_numpy(a).to(dev) for a in case.launch], torch.from_numpy(case.amp).to(dev), n_global)
    nbins = 64

    def one():
        # static grid: a clean beam has no stragglers, and CTAs that retire let the exchange kernels in
        return rd.trace_sharded(field, None, staged=staged, nbins=nbins, hist_kw=hist_kw, gather="root", comm=comm,
                                chunks=args.config5_chunks, refill=False, timing=True, **kw)

    for _ in range(2):
        res, hist, gathered, tm = one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = max(3, min(args.steps, 5))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    tms = []
    e0.record()
    for _ in range(reps):
        res, hist, gathered, tm = one()
        tms.append(tm)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    parts = np.array([[t.ms()[k] for k in ("kernel_ms", "total_ms", "exchange_exposed_ms")] for t in tms]).mean(0)
    v = torch.tensor([e0.elapsed_time(e1) / reps, parts[0], parts[1], parts[2]], dtype=torch.float64, device=dev)
    vmin = v.clone()
    work = res.nsteps.sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_pass, ms_kernel, ms_total, ms_wait = (float(x) for x in v.tolist())
    ms_kernel_min = float(vmin[1])
    # what the exchange adds to the pass = what is left after the SLOWEST rank's last trace kernel has finished
    # (a rank that finishes tracing early waits for that rank in the all-reduce: that wait is load skew, not exchange)
    ms_exposed = max(ms_pass - ms_kernel, 0.0)
    out = None
    if rank == 0:
        # spot check of the gathered result (untimed): 4096 global rays, their step counts and end states
        # in GLOBAL order against the oracle on the same launch states
        import oracle
        sel = np.unique(np.linspace(0, n_global - 1, 4096).astype(np.int64))
        sub = synth.beam_case(nx=201, nrays=n_global, seed=0, index=sel, device_launch=True)
        w = oracle.trace_batch(*sub.launch, case.L, case.st, case.dx, case.dy, nf.nx, nf.ny, np.ascontiguousarray(nf.ax),
                               np.ascontiguousarray(nf.ay), nthreads=4)
        isel = torch.from_numpy(sel).to(dev)
        same = bool(np.array_equal(gathered.nsteps[isel].cpu().numpy(), w["nsteps"])) and \
            gathered.end_state[:, isel].cpu().numpy().tobytes() == w["end"].tobytes()
        out = {
            "workload": f"config5: {n_global}-ray seven-beam fan on a 201x201 grid sharded over {world} GPU(s) "
                        f"({args.config5_rays} rays each, 256-ray blocks round-robin), endpoints + {nbins}-bin deposition "
                        f"histogram; NCCL all-reduce of the histogram; step counts + end states of ALL rays gathered in "
                        f"global ray order on rank 0 (36 B/ray), {args.config5_chunks} chunks, exchange of chunk k under the trace "
                        f"of chunk k+1",
            "value": float(work.item()) / (ms_pass * 1e-3), "unit": UNIT, "ms_per_step": ms_pass, "steps": reps,
            "ray_steps_per_step": float(work.item()), "kernel_ms": ms_kernel, "kernel_ms_fastest_rank": ms_kernel_min,
            "exchange_exposed_ms": ms_exposed, "exchange_frac_of_pass": ms_exposed / ms_pass if ms_pass else None,
            "wait_after_own_trace_ms_max_rank": ms_wait,
            "note": "kernel_ms = the slowest rank's trace (the root: NCCL's receive and the unshuffle kernels of earlier chunks share "
                    "its SMs); exchange_exposed_ms = pass - kernel_ms, the exchange left on the critical path; the wait of the "
                    "fastest rank for the slowest one in the all-reduce is load skew, reported separately",
            "gathered_bytes_on_root": int(n_global) * 36, "hist_sum": float(hist.sum().item()),
            "transport": "stb_comm_* (NCCL via the C ABI)" if comm is not None else "single rank: no exchange",
            "gathered_spot_check": {"result": "ok" if same else "MISMATCH", "rays": int(len(sel)),
                                    "against": "oracle.trace_batch on the same launch states, global ray order, untimed"},
        }
    field.close()
    del staged, gathered
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from rayrtrace_b200 import synth
    # this rank's interleaved shard of the N*nrays global fan (interleave for balance)
    case = synth.beam_case(nx=args.nx, nrays=args.nrays * world, seed=0, L=args.cap or None)
    if world > 1:
        sl = slice(rank, None, world)
        case.launch = tuple(np.ascontiguousarray(a[sl]) for a in case.launch)
        case.amp = case.amp[sl]
        case.nrays = len(case.launch[0])

    # CPU baseline first (forks worker processes: must happen before CUDA is initialised)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(case)
        rays = calibrate_cpu(arm, args.cpu_seconds / 2)
        s1, t1, n1 = arm.step(rays)
        s2, t2, _ = arm.step(rays)
        arm.close()
        s, t = (s1, t1) if s1 / t1 > s2 / t2 else (s2, t2)
        cpu = {"value": s / t, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
               "sample": f"{n1} rays drawn (seed 1234) from the same {case.nrays}-ray workload = {s} ray-steps, best of 2, "
                         f"one process per core calling trace() per ray with prefaulted buffers"}

    import torch
    import torch.distributed as dist
    from rayrtrace_b200 import _cabi, dist as rd, sidtrace
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = rd.Comm.from_torch(dev)

    nf = case.nfield
    field = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny)
    n, L = case.nrays, case.L
    d_in = [torch.from_numpy(a).to(dev) for a in case.launch]
    tiled = args.layout == "tiled"
    traj = torch.empty(((n + 255) // 256, L, 256, 4) if tiled else (n, L, 4), dtype=torch.float64, device=dev)
    out = sidtrace.BatchResult(nsteps=torch.empty(n, dtype=torch.int32, device=dev),
                               end_state=torch.empty((4, n), dtype=torch.float64, device=dev), traj=traj)
    kw = dict(L=L, st=case.st, dx=case.dx, dy=case.dy, tiled=tiled)

    def step():
        sidtrace.trace_batch(field, *d_in, traj=traj, out=out, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)          # the contract asks for >= 3 warm-up steps; the line reports what ran
    for _ in range(warmup):
        step()
    barrier()
    steps_per_launch = int(out.nsteps.sum().item())

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _cabi.launch_count()
    barrier()
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for a, b in ev:
        a.record()
        step()
        b.record()
    t_all1.record()
    barrier()
    launches = _cabi.launch_count() - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    clk = clocks.stop() if rank == 0 else None

    # max over ranks (device time), sum of work
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    work = torch.tensor([float(steps_per_launch)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    total_ms = float(tm.item())
    global_steps_per_step = float(work.item())
    value = global_steps_per_step * args.steps / (total_ms * 1e-3)

    # ---- untimed: the trajectories the timed kernel wrote, against the oracle ----------------------
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_digest(case, traj, out.nsteps, tiled)

    # ---- end to end through the public API with HOST buffers (pinned): H2D launch states,
    # kernel (full trajectories into HBM, kept resident for the device-side consumers),
    # D2H step counts + end states.
    h_in = [torch.from_numpy(a).pin_memory().numpy() for a in case.launch]
    h_out = sidtrace.BatchResult(nsteps=torch.empty(n, dtype=torch.int32).pin_memory().numpy(),
                                 end_state=torch.empty((4, n), dtype=torch.float64).pin_memory().numpy(), traj=traj)
    for _ in range(2):
        sidtrace.trace_batch(field, *h_in, traj=traj, out=h_out, **kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sidtrace.trace_batch(field, *h_in, traj=traj, out=h_out, **kw)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = global_steps_per_step * args.steps / float(e2e_t.item())
    assert int(h_out.nsteps.astype(np.int64).sum()) == steps_per_launch
    h2d = 4 * n * 8
    d2h = n * 4 + 4 * n * 8

    # ---- for transparency: the same metric when EVERY state must reach host memory (packed
    # export, PCIe-bound at 32 B per ray-step); bounded sample, rank 0 only.
    traj_host = None
    if rank == 0 and not args.no_traj_export:
        from rayrtrace_b200 import export
        m = min(n, 62_500)
        sel = np.linspace(0, n - 1, m).astype(np.int64)
        sub = [np.ascontiguousarray(a[sel]) for a in case.launch]
        out_pin = sidtrace.pinned_empty((int(m * (steps_per_launch / n) * 1.3) + 1024, 4))
        export.trace_to_host(field, *sub, L=L, st=case.st, dx=case.dx, dy=case.dy, out=out_pin)     # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        off, _, _, _ = export.trace_to_host(field, *sub, L=L, st=case.st, dx=case.dx, dy=case.dy, out=out_pin)
        dt = time.perf_counter() - t0
        traj_host = {"value": float(off[-1]) / dt, "unit": UNIT, "d2h_bytes": int(off[-1]) * 32, "sample_rays": int(m),
                     "note": "export.trace_to_host: chunked trace, pack, D2H of every kept state into pinned host memory, "
                             "copies overlapped with tracing on a second stream; PCIe-bound"}
        del out_pin

    # ---- the O(rays)-output mode of the same workload against the measured FP64 peak, and the
    # two-pass exactly-sized arena (rank 0; a few ms each)
    endpoint = exact = None
    if rank == 0:
        import ctypes
        tf = ctypes.c_double()
        _cabi.check(_cabi.lib.stb_measure_fp64_peak(ctypes.byref(tf)))
        bw = ctypes.c_double()
        _cabi.check(_cabi.lib.stb_measure_store_bw(8 << 30, ctypes.byref(bw)))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        cells = sidtrace.pick_cells(d_in[2], d_in[3])     # the fan climbs in y: the y-major table (picked outside the timed region)
        kw0 = dict(kw, tiled=False, refill=False, cells=cells)
        o0 = sidtrace.BatchResult(nsteps=out.nsteps, end_state=out.end_state)
        for _ in range(3):
            sidtrace.trace_batch(field, *d_in, out=o0, **kw0)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            sidtrace.trace_batch(field, *d_in, out=o0, **kw0)
        e1.record()
        torch.cuda.synchronize()
        ep = steps_per_launch * args.steps / (e0.elapsed_time(e1) * 1e-3)
        endpoint = {"value": ep, "unit": UNIT, "ms_per_launch": e0.elapsed_time(e1) / args.steps,
                    "fp64_peak_tflops_measured": tf.value, "store_bw_gbs_measured": bw.value,
                    "fp64_frac": ep * FLOPS_PER_RAY_STEP / (tf.value * 1e12) if tf.value else None,
                    "cells": cells + "-major",
                    "dispatch_ceiling": "7.2e11 ray-steps/s with every gather an L1 hit (profiles/r02/dual_ray_endpoint.txt): an FP64 "
                                        "instruction holds its scheduler's dispatch port two cycles, 34.5 + 14 cycles per ray-step",
                    "note": "step counts + end states only (static grid, two rays per thread, hoisted per-cell products); 16 algorithmic "
                            "flops per ray-step (SURVEY 8d) over a dependent-free DFMA peak measured in this run"}
        del traj
        out.traj = h_out.traj = None
        torch.cuda.empty_cache()
        kx = dict(L=L, st=case.st, dx=case.dx, dy=case.dy, refill=False, cells=cells)   # a clean, ordered beam: pass 1 on the static grid
        r = None
        for _ in range(2):
            r = sidtrace.trace_exact(field, *d_in, **kx)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(2, args.steps // 2)):
            r = sidtrace.trace_exact(field, *d_in, **kx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / max(2, args.steps // 2)
        exact = {"value": steps_per_launch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                 "arena_bytes": int(r.plan.arena_bytes), "cap_sized_arena_bytes": int(((n + 255) // 256) * L * 256 * 32),
                 "frac_of_hbm_peak": steps_per_launch * 32 / (ms * 1e-3) / 1e9 / load_peaks()[0],
                 "note": "sidtrace.trace_exact(refill=False): pass 1 (step counts, no trajectories; static grid, two rays per thread) + stb_traj_plan (sort by length, "
                         "exact tile sizes) + pass 2 (trajectories into the packed arena); the whole of it timed"}
        del r
        torch.cuda.empty_cache()

    extra = {}
    if args.extra and rank == 0:
        # the adaptive-RK extension (never part of the headline): 10^6 ordered rays, 1001^2 spline medium
        from rayrtrace_b200 import physics
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        xs_m = np.linspace(-10, 10, 1001)
        XX, YY = np.meshgrid(xs_m, xs_m)
        med = physics.Medium(0.6 * np.exp(-2.5 * (XX ** 2 + YY ** 2) / 7.85 ** 2), xs_m, xs_m)
        rx0 = torch.linspace(-5, 5, 1_000_000, dtype=torch.float64, device=dev)
        ry0 = torch.full_like(rx0, -9.5)
        rn = physics.refractive_index(med, rx0, ry0)
        rpx, rpy = (rn * torch.sin(0.03 * rx0)).contiguous(), (rn * torch.cos(0.03 * rx0)).contiguous()
        rres = physics.RkResult()
        for _ in range(3):
            physics.trace_rk(med, rx0, ry0, rpx, rpy, rtol=1e-8, atol_r=1e-8, atol_p=1e-8, hmax=0.4, out=rres)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            physics.trace_rk(med, rx0, ry0, rpx, rpy, rtol=1e-8, atol_r=1e-8, atol_p=1e-8, hmax=0.4, out=rres)
        e1.record()
        torch.cuda.synchronize()
        att = int(rres.nacc.sum() + rres.nrej.sum())
        extra["rk_extension"] = {"attempts_per_s": att * 5 / (e0.elapsed_time(e1) * 1e-3), "ms_per_launch": e0.elapsed_time(e1) / 5,
                                 "workload": "1e6 ordered rays, 1001^2 B-spline medium, O mode, Dormand-Prince 5(4), rtol 1e-8",
                                 "note": "extension beyond the reference (SURVEY 8 f-4); not parity, not the headline"}
        med.close()

    # ---- application level (rank 0): the reference's driver shape, wall clock --------------------------
    app = None
    if rank == 0 and not args.no_app:
        app = run_app(_cabi)

    # ---- BASELINE config 5: the sharded path with its exchange, every rank ---------------------------
    field.close()
    del d_in
    out.traj = None
    torch.cuda.empty_cache()
    config5 = None
    if not args.no_config5:
        config5 = run_config5(args, rank, world, local_rank, dev, comm)

    if world > 1:
        dist.barrier()
        if comm is not None:
            comm.close()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peak, peak_src = load_peaks()
    avg_kernel_ms = float(np.mean(kern_ms))
    achieved = BYTES_PER_RAY_STEP * steps_per_launch / (avg_kernel_ms * 1e-3) / 1e9
    tr = load_traffic() if world == 1 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "warmup_requested": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": make_config(args, L),
        "ray_steps_per_gpu_per_step": steps_per_launch,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (tr or {}).get("dram_bytes_per_launch"),
                     "traffic_source": (tr or {}).get("source") if tr else
                     ("not attached at N > 1: the ncu capture is a single-GPU profiling run" if world > 1 else None),
                     "kernel": "stb::trace_kernel<TRAJ=1,EXTRAS=0,HIST=0,DIVC=1,REFILL=0,TILED=%d,CROSS=0,STRAY=0,LOOKUP=0,BLOCK=%d,HOIST=0>" % (int(tiled), 256 if tiled else 128),
                     "algorithmic_bytes_per_launch": BYTES_PER_RAY_STEP * steps_per_launch,
                     "avg_launch_ms": avg_kernel_ms, "peak_source": peak_src,
                     "frac_vs_store_peak": (achieved / endpoint["store_bw_gbs_measured"]) if endpoint else None},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "host launch states in (pinned) -> kernel -> per-ray step counts + end states out; "
                        "trajectories stay resident in HBM"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "parity_check": parity,
        "endpoint_mode": endpoint,
        "exact_arena": exact,
        "config5": config5,
        "app": app,
    }
    if traj_host:
        line["e2e_traj_to_host"] = traj_host
    if extra:
        line["extra"] = extra
    emit(line)
    return 0


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else any library prints
    to fd 1 (e.g. NCCL's version banner) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
