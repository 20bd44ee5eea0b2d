"""The adaptive-RK CUDA kernel (include/sidtrace_b200_rk.h) against its CPU checker, through the C ABI.

EXTENSION, PARITY UNPINNED: no reference output exists for this mode; the checker
(oracle/rk_oracle.c) is pinned by closed forms and SciPy in tests/test_rk_oracle.py.  Bar here:
BIT-EXACT on everything a ray carries (states, step counts, statuses, tau / phi / sigma, the
accepted-state trajectories, the prefiltered spline table); the deposition histogram within 1e-9
relative L1 (FP64 atomics and exp() differ in the last bits).  Plus closed-form and conservation
properties at sizes the CPU checker does not reach.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INTERPS = ["bilinear", "catmull-rom", "bspline"]


@pytest.fixture(scope="module")
def ph():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from rayrtrace_b200 import physics
    return physics


@pytest.fixture(scope="module")
def rk():
    from oracle import rk
    rk.lib()
    return rk


def _tables(n=201, peak=0.6, seed=0, with_alpha=False):
    xs = np.linspace(-10, 10, n)
    XX, YY = np.meshgrid(xs, xs)
    rng = np.random.default_rng(seed)
    X = peak * np.exp(-((XX - 0.4) ** 2 + 1.3 * YY ** 2) / 14.0) * (1 + 0.05 * np.sin(1.7 * XX) * np.cos(1.1 * YY))
    X += 1e-3 * rng.random(X.shape)          # table noise: exercises every stencil value
    alpha = 0.02 * np.exp(-(XX ** 2 + YY ** 2) / 9.0) if with_alpha else None
    return xs, X, alpha


def _beam(n, seed=1):
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-6, 6, n)
    y0 = np.full(n, -9.5) + rng.uniform(0, 0.3, n)
    th = rng.uniform(-0.5, 0.5, n)
    return x0, y0, np.sin(th), np.cos(th)


def _gpu(ph, med, x0, y0, px, py, **kw):
    import torch
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for k in ("xscale", "yscale", "amp"):
        if k in kw:
            kw[k] = dev(kw[k])
    res = ph.trace_rk(med, dev(x0), dev(y0), dev(px), dev(py), **kw)
    torch.cuda.synchronize()
    return res


def _same(res, want, keys=("end_state", "tau", "phi", "sig", "n2_min", "ham", "nacc", "nrej", "status", "turn_index")):
    for k in keys:
        got = getattr(res, k).cpu().numpy()
        assert got.tobytes() == want[k].tobytes(), (k, np.flatnonzero((got != want[k]).reshape(-1))[:5])
    assert np.array_equal(res.turn_state.cpu().numpy(), want["turn_state"], equal_nan=True)


INTERP_ID = {"bilinear": 0, "catmull-rom": 1, "bspline": 2}


def test_device_prefilter_is_bitwise_the_checkers(ph, rk):
    xs, X, alpha = _tables(151, with_alpha=True)
    X = X[:, :131].copy()                       # nx != ny
    alpha = alpha[:, :131].copy()
    med = ph.Medium(X, xs[:131], xs, alpha=alpha)
    assert med.table("X").tobytes() == rk.prefilter(X).tobytes()
    assert med.table("alpha").tobytes() == rk.prefilter(alpha).tobytes()
    raw = ph.Medium(X, xs[:131], xs, interp="catmull-rom")
    assert raw.table("X").tobytes() == X.tobytes()
    for shape in ((2, 2), (2, 5), (3, 2)):
        G = np.random.default_rng(4).random(shape)
        t = ph.Medium(G, np.arange(shape[1]) * 1.0, np.arange(shape[0]) * 1.0).table()
        assert t.tobytes() == rk.prefilter(G).tobytes()


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("absorb", [False, True])
def test_every_variant_bitwise(ph, rk, interp, mode, absorb):
    xs, X, alpha = _tables(with_alpha=absorb, peak=0.6 if mode == "O" else 0.3)
    d = xs[1] - xs[0]
    n = 3000
    x0, y0, px, py = _beam(n)
    rng = np.random.default_rng(7)
    xscale = rng.uniform(0.2, 2.2, n)           # under-dense ... reflected
    yscale = rng.uniform(0.8, 1.2, n)
    amp = rng.uniform(0.5, 1.5, n)
    kw = dict(rtol=1e-8, atol_r=1e-8, atol_p=1e-8, h0=0.05, hmin=1e-10, hmax=0.4, max_steps=5000)
    med = ph.Medium(X, xs, xs, alpha=alpha, mode=mode, Y0=0.45, R0=55.0, interp=interp)
    res = _gpu(ph, med, x0, y0, px, py, xscale=xscale, yscale=yscale, amp=amp, L=48, aux=True,
               nbins=32 if absorb else 0, r_lcs=7.85, rho_max=1.5, **kw)
    om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, alpha=alpha, mode=mode, Y0=0.45, R0=55.0, interp=INTERP_ID[interp])
    want = rk.trace(om, x0, y0, px, py, xscale=xscale, yscale=yscale, amp=amp, L=48, aux=True,
                    nbins=32 if absorb else 0, r_lcs=7.85, rho_max=1.5, threads=8, **kw)
    _same(res, want)
    assert res.traj.cpu().numpy().tobytes() == want["traj"].tobytes()
    assert res.aux.cpu().numpy().tobytes() == want["aux"].tobytes()
    assert len(set(want["status"].tolist())) >= 1 and want["nacc"].min() > 3
    if absorb:
        got = res.hist.cpu().numpy()
        assert np.abs(got - want["hist"]).sum() <= 1e-9 * np.abs(want["hist"]).sum()
        assert want["hist"].sum() > 0


def test_sweep_with_cutoffs_resonances_and_caps_bitwise(ph, rk):
    """2*10^5 rays, X mode, density scale 0..6 and field scale 0.5..1.5: under-dense, reflected at
    the right-hand cutoff, collapsing on the upper-hybrid layer, capped; every status occurs."""
    xs, X, _ = _tables(301, peak=0.5, seed=3)
    d = xs[1] - xs[0]
    n = 200_000
    x0, y0, px, py = _beam(n, seed=5)
    rng = np.random.default_rng(11)
    xscale = rng.uniform(0.0, 6.0, n)
    yscale = rng.uniform(0.5, 1.5, n)
    # a few rays start inside the dense core (beyond the upper-hybrid layer) or outside the table
    x0[:2000] = rng.uniform(-1, 1, 2000); y0[:2000] = rng.uniform(-1, 1, 2000); xscale[:2000] = rng.uniform(2.0, 3.4, 2000)
    x0[2000:2100] = 10.5
    kw = dict(rtol=1e-7, atol_r=1e-7, atol_p=1e-7, h0=0.05, hmin=1e-9, hmax=0.4, max_steps=400, tau_max=60.0)
    med = ph.Medium(X, xs, xs, mode="X", Y0=0.45, R0=55.0)
    res = _gpu(ph, med, x0, y0, px, py, xscale=xscale, yscale=yscale, **kw)
    om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, mode="X", Y0=0.45, R0=55.0)
    want = rk.trace(om, x0, y0, px, py, xscale=xscale, yscale=yscale, threads=16, **kw)
    _same(res, want)
    assert set(want["status"].tolist()) >= {rk.LEFT_DOMAIN, rk.MAX_STEPS, rk.STEP_UNDERFLOW, rk.START_OUTSIDE}


@pytest.mark.parametrize("stop", [(5.0, "y"), (3.0, "x")])
@pytest.mark.parametrize("absorb", [False, True])
def test_stop_line_bitwise_and_on_the_line(ph, rk, stop, absorb):
    """Receiver line: rays end ON the line (regula falsi on the step length); the landing trials, the
    landed states and the optical path are the checker's bit for bit; rays that never reach the line
    end as usual."""
    xs, X, alpha = _tables(with_alpha=absorb, peak=0.5)
    n = 4000
    x0, y0, px, py = _beam(n, seed=9)
    rng = np.random.default_rng(13)
    xscale = rng.uniform(0.2, 2.5, n)
    kw = dict(rtol=1e-9, atol_r=1e-9, atol_p=1e-9, h0=0.05, hmin=1e-10, hmax=0.4, max_steps=5000, stop=stop)
    med = ph.Medium(X, xs, xs, alpha=alpha)
    res = _gpu(ph, med, x0, y0, px, py, xscale=xscale, L=32, aux=True, nbins=16 if absorb else 0, r_lcs=7.85, rho_max=1.5, **kw)
    om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, alpha=alpha)
    want = rk.trace(om, x0, y0, px, py, xscale=xscale, L=32, aux=True, nbins=16 if absorb else 0, r_lcs=7.85, rho_max=1.5,
                    threads=8, **kw)
    _same(res, want)
    assert res.traj.cpu().numpy().tobytes() == want["traj"].tobytes()
    landed = want["status"] == rk.STOP_LINE
    assert 0.3 * n < landed.sum() < n                       # some rays are reflected or leave sideways first
    coord = want["end_state"][1 if stop[1] == "y" else 0]
    assert np.abs(coord[landed] - stop[0]).max() <= 1e-12
    if absorb:
        got = res.hist.cpu().numpy()
        assert np.abs(got - want["hist"]).sum() <= 1e-9 * np.abs(want["hist"]).sum()


def test_fuzz_random_media_and_controls_bitwise(ph, rk):
    """Random table shapes and spacings, interpolants, modes, tolerances, step limits, scales, stop
    lines, caps and batch sizes: every per-ray output bit-identical to the checker."""
    import os
    rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "7")))
    for it in range(int(os.environ.get("FUZZ_ITERS", "40"))):
        nx, ny = int(rng.integers(2, 160)), int(rng.integers(2, 160))
        xs = np.linspace(-rng.uniform(2, 12), rng.uniform(2, 12), nx)
        ys = np.linspace(-rng.uniform(2, 12), rng.uniform(2, 12), ny)
        XX, YY = np.meshgrid(xs, ys)
        X = rng.uniform(0, 1.3) * np.exp(-(XX ** 2 + YY ** 2) / rng.uniform(4, 40)) + 0.02 * rng.random(XX.shape)
        absorb = bool(rng.random() < 0.4)
        alpha = 0.05 * rng.random(XX.shape) if absorb else None
        mode = "X" if rng.random() < 0.5 else "O"
        interp = INTERPS[int(rng.integers(0, 3))]
        n = int(rng.choice([1, 2, 31, 33, 100, 257, 1500]))
        x0 = rng.uniform(xs[0] - 0.5, xs[-1] + 0.5, n)
        y0 = rng.uniform(ys[0] - 0.5, ys[-1] + 0.5, n)
        th = rng.uniform(0, 2 * np.pi, n)
        pm = rng.uniform(0.1, 1.2, n)
        if n > 4:
            x0[rng.integers(0, n)] = rng.choice([np.nan, 1e30, -1e30, xs[0], xs[-1]])
        tol = float(10.0 ** rng.uniform(-11, -4))
        cell = min(xs[1] - xs[0], ys[1] - ys[0])
        h0 = float(cell * rng.uniform(0.05, 2.0))
        kw = dict(rtol=tol, atol_r=tol, atol_p=tol, h0=h0, hmin=float(h0 * 10.0 ** rng.uniform(-9, -2)),
                  hmax=float(h0 * rng.uniform(1, 20)), max_steps=int(rng.choice([0, 1, 7, 300])),
                  tau_max=float(rng.choice([np.inf, 0.37, 5.0, 40.0])))
        if rng.random() < 0.5:
            kw["stop"] = (float(rng.uniform(ys[0], ys[-1])), "y") if rng.random() < 0.5 else (float(rng.uniform(xs[0], xs[-1])), "x")
        per_ray = {}
        if rng.random() < 0.5:
            per_ray["xscale"] = rng.uniform(0, 3, n)
        if rng.random() < 0.5:
            per_ray["yscale"] = rng.uniform(0.3, 2, n)
        if rng.random() < 0.5:
            per_ray["amp"] = rng.uniform(0, 2, n)
        L = int(rng.choice([0, 1, 5, 40]))
        hist = dict(nbins=int(rng.choice([1, 7, 64])), r_lcs=float(rng.uniform(2, 9)), rho_max=float(rng.uniform(0.5, 2)),
                    xaxis=float(rng.uniform(-1, 1)), yaxis=float(rng.uniform(-1, 1))) if absorb and rng.random() < 0.7 else {}
        med = ph.Medium(X, xs, ys, alpha=alpha, mode=mode, Y0=float(rng.uniform(0.2, 0.7)), R0=float(rng.uniform(20, 80)), interp=interp)
        res = _gpu(ph, med, x0, y0, pm * np.cos(th), pm * np.sin(th), L=L, aux=bool(L and rng.random() < 0.5), **kw, **per_ray, **hist)
        om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, alpha=alpha, mode=mode, Y0=med.Y0, R0=med.R0, interp=INTERP_ID[interp])
        want = rk.trace(om, x0, y0, pm * np.cos(th), pm * np.sin(th), L=L, aux=res.aux is not None, threads=4, **kw, **per_ray, **hist)
        ctx = (it, nx, ny, n, mode, interp, absorb, kw, L)
        for k in ("nacc", "nrej", "status", "tau", "phi", "sig", "n2_min", "end_state", "ham", "turn_index", "turn_state"):
            got = getattr(res, k).cpu().numpy()
            # NaN launch states propagate NaNs whose payload/sign may differ between CPU and GPU: compare as values
            assert np.array_equal(got, want[k], equal_nan=True), (k, ctx)
        if L:
            assert np.array_equal(res.traj.cpu().numpy(), want["traj"], equal_nan=True), ctx
        if hist:
            got, w = res.hist.cpu().numpy(), want["hist"]
            assert np.abs(got - w).sum() <= 1e-9 * max(np.abs(w).sum(), 1e-300), ctx
        med.close()


def test_results_do_not_depend_on_the_schedule(ph):
    """Warp tiles (32), the default (24), lane refill (1): which warp traces a ray, and next to which
    other rays, must not change a bit of its result -- including stragglers 40x longer than the rest."""
    import torch
    xs, X, alpha = _tables(with_alpha=True)
    X[:60] = 0.0                                          # a vacuum strip along the launch line
    n = 50_000
    x0, y0, px, py = _beam(n, seed=3)
    px[::97] = 0.0; py[::97] = 0.0                        # these never move: they run into the step cap
    med = ph.Medium(X, xs, xs, alpha=alpha, interp="catmull-rom")   # local interpolant: the strip stays exactly 0
    outs = []
    for kw in (dict(refill_min=32), dict(), dict(lane_refill=True), dict(refill_min=7)):
        r = _gpu(ph, med, x0, y0, px, py, rtol=1e-8, atol_r=1e-8, atol_p=1e-8, hmax=0.4, max_steps=4000, nbins=32,
                 r_lcs=7.85, rho_max=1.5, **kw)
        outs.append(r)
    assert int((outs[0].status == ph.MAX_STEPS).sum()) == len(range(0, n, 97))
    for r in outs[1:]:
        for k in ("end_state", "tau", "phi", "sig", "n2_min", "ham", "nacc", "nrej", "status", "turn_index"):
            assert torch.equal(getattr(r, k), getattr(outs[0], k)), k
        assert torch.equal(torch.nan_to_num(r.turn_state, nan=-7.0), torch.nan_to_num(outs[0].turn_state, nan=-7.0))
        torch.testing.assert_close(r.hist, outs[0].hist, rtol=1e-11, atol=0)


def test_sorted_tracing_returns_the_same_results_in_the_callers_order(ph):
    import torch
    xs, X, alpha = _tables(with_alpha=True)
    n = 30_000
    x0, y0, px, py = _beam(n, seed=17)                    # random positions and angles: an incoherent set
    x0[5] = np.nan
    med = ph.Medium(X, xs, xs, alpha=alpha)
    kw = dict(rtol=1e-8, atol_r=1e-8, atol_p=1e-8, hmax=0.4, L=8, aux=True, nbins=16, r_lcs=7.85, rho_max=1.5)
    a = _gpu(ph, med, x0, y0, px, py, **kw)
    b = _gpu(ph, med, x0, y0, px, py, sort_rays=True, **kw)
    for k in ("end_state", "tau", "phi", "sig", "n2_min", "ham", "nacc", "nrej", "status", "traj", "aux", "turn_index", "turn_state"):
        assert torch.equal(torch.nan_to_num(getattr(a, k).double(), nan=-7.0), torch.nan_to_num(getattr(b, k).double(), nan=-7.0)), k
    torch.testing.assert_close(a.hist, b.hist, rtol=1e-11, atol=0)
    order = ph.coherent_order(med, *[torch.from_numpy(v).cuda() for v in (x0, y0, px, py)])
    assert sorted(order.tolist()) == list(range(n))


def test_whole_warps_ending_at_launch_and_tiny_batches(ph, rk):
    xs, X, _ = _tables(101)
    d = xs[1] - xs[0]
    med = ph.Medium(X, xs, xs)
    om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy)
    d = med.dx
    for n in (1, 31, 33, 1000):
        x0 = np.full(n, 12.0); x0[n // 2:] = -3.0          # first half starts outside the table
        y0 = np.full(n, -9.0)
        res = _gpu(ph, med, x0, y0, np.zeros(n), np.ones(n), hmax=0.4)
        want = rk.trace(om, x0, y0, np.zeros(n), np.ones(n), hmax=0.4, h0=0.5 * d)
        _same(res, want)
    # every ray ends at launch (the refill loop must keep claiming)
    n = 5000
    res = _gpu(ph, med, np.full(n, 12.0), np.zeros(n), np.ones(n), np.zeros(n))
    assert (res.status.cpu().numpy() == ph.START_OUTSIDE).all() and (res.nacc.cpu().numpy() == 0).all()
    import torch
    e = torch.empty(0, dtype=torch.float64, device="cuda")
    assert ph.trace_rk(med, e, e, e, e).end_state.shape == (4, 0)


def test_refractive_index_helper_and_on_shell_launch(ph, rk):
    import torch
    xs, X, _ = _tables(201, peak=0.3)
    med = ph.Medium(X, xs, xs, mode="X", Y0=0.45, R0=55.0)
    om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, mode="X", Y0=0.45, R0=55.0)
    x0 = torch.linspace(-5, 5, 257, dtype=torch.float64, device="cuda")
    y0 = torch.full_like(x0, -9.5)
    N = ph.refractive_index(med, x0, y0)
    want = np.sqrt([om.eval(float(x), -9.5)[3] for x in x0.cpu()])
    np.testing.assert_array_equal(N.cpu().numpy(), want)
    res = ph.trace_rk(med, x0, y0, torch.zeros_like(x0), N, rtol=1e-10, atol_r=1e-10, atol_p=1e-10)
    assert float(res.ham.abs().max()) < 2e-6 and bool((res.status == ph.LEFT_DOMAIN).all())   # noisy table


def test_million_ray_properties(ph):
    """Sizes the CPU checker does not reach: conservation, deposited power, closed-form ramp."""
    import torch
    n = 1_000_000
    xs = np.linspace(-10, 10, 401)
    XX, YY = np.meshgrid(xs, xs)
    X = 0.7 * np.exp(-(XX ** 2 + YY ** 2) / 16.0)
    alpha = np.full_like(X, 0.02)
    med = ph.Medium(X, xs, xs, alpha=alpha)
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = (torch.rand(n, generator=g, dtype=torch.float64, device="cuda") - 0.5) * 12
    y0 = torch.full_like(x0, -9.5)
    th = (torch.rand(n, generator=g, dtype=torch.float64, device="cuda") - 0.5) * 0.8
    N = ph.refractive_index(med, x0, y0)
    amp = 0.5 + torch.rand(n, generator=g, dtype=torch.float64, device="cuda")
    res = ph.trace_rk(med, x0, y0, N * torch.sin(th), N * torch.cos(th), amp=amp, rtol=1e-9, atol_r=1e-9, atol_p=1e-9,
                      nbins=64, r_lcs=7.85, rho_max=2.5)
    torch.cuda.synchronize()
    assert bool((res.status == ph.LEFT_DOMAIN).all())
    assert float(res.ham.abs().max()) < 1e-6
    torch.testing.assert_close(res.sig, 0.02 * res.tau, rtol=1e-12, atol=0)
    absorbed = (amp * amp * (1 - torch.exp(-res.sig))).sum()
    assert float(res.hist.sum()) == pytest.approx(float(absorbed), rel=1e-10)

    # linear ramp: exact parabola for every ray
    Xr = np.tile(0.4 + 0.04 * xs, (401, 1))
    ramp = ph.Medium(Xr, xs, xs, interp="catmull-rom")
    x0 = torch.full((n,), -9.5, dtype=torch.float64, device="cuda")
    y0 = (torch.rand(n, generator=g, dtype=torch.float64, device="cuda") - 0.5) * 10
    th = (torch.rand(n, generator=g, dtype=torch.float64, device="cuda") - 0.5) * 1.0
    N0 = np.sqrt(1 - (0.4 + 0.04 * -9.5))
    px, py = N0 * torch.cos(th), N0 * torch.sin(th)
    T = 6.0
    res = ph.trace_rk(ramp, x0, y0, px, py, tau_max=T, hmax=0.5)
    torch.cuda.synchronize()
    assert bool((res.status == ph.TAU_MAX).all())
    want = torch.stack([x0 + px * T - 0.25 * 0.04 * T * T, y0 + py * T, px - 0.5 * 0.04 * T, py])
    torch.testing.assert_close(res.end_state, want, rtol=0, atol=5e-12)


def test_argument_errors(ph):
    import torch
    from rayrtrace_b200._cabi import SidtraceError
    xs, X, _ = _tables(51)
    med = ph.Medium(X, xs, xs)
    z = torch.zeros(4, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        ph.trace_rk(med, z, z, z, z, nbins=8)                       # histogram without alpha
    with pytest.raises(SidtraceError):
        ph.trace_rk(med, z, z, z, z, atol_r=0.0)
    with pytest.raises(SidtraceError):
        ph.trace_rk(med, z, z, z, z, h0=1.0, hmax=0.5)
    with pytest.raises(TypeError):
        ph.trace_rk(med, z.cpu(), z, z, z)
    with pytest.raises(ValueError):
        ph.Medium(X, xs ** 3, xs)                                   # non-uniform grid
    with pytest.raises(SidtraceError):
        ph.Medium(X, xs, xs, mode="X", R0=0.0)


def test_plain_c_caller_with_host_pointers(ph, tmp_path):
    """tests/c/rk_abi.c: no Python, no torch -- STB_RK_F_HOST_IO, closed-form checks."""
    import os
    import subprocess
    from rayrtrace_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "rk_abi"
    lib = _cabi.LIB_PATH
    subprocess.check_call([os.environ.get("CC", "gcc"), "-O1", "-std=c11", os.path.join(root, "tests", "c", "rk_abi.c"),
                           lib, "-lm", "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout and "all ok" in r.stdout


def test_beam_receiver_phases_parity_path_converges_to_rk(ph):
    """Closing the loop with the reference's workflow: one FT-2 beam, launch set of Beam.rays
    (beam.py:278-331), receiver line on the far side.  Per ray, the parity path's eikonal phase at the
    line (fused crossing output = Ray.addpoint + Ray.ph, ray.py:113-224) approaches the RK phase at
    first order in the grid spacing; the RK phases themselves move several times less between the grids."""
    import torch
    from _util import ft2_density
    from rayrtrace_b200.beam import Beam
    from rayrtrace_b200.nfield import Nfield
    x_ant = 2.3
    y_ant = -float(np.sqrt(100 - x_ant ** 2))
    worst, rk_phase, masks = [], [], []
    for nx in (201, 401):
        n, x, y = ft2_density(nx, 4e13)
        nf = Nfield(n, x=x, y=y, freq=136.268e9)
        b = Beam(x0=x_ant, y0=y_ant, kx0=0, ky0=1)
        b.rays(rnum=101, dist=3, ampbnd=1e-3, n_field=nf, receiver=(-y_ant, "y"))
        cs = b.bundle.result.cross_state
        cs = cs.cpu().numpy() if hasattr(cs, "cpu") else np.asarray(cs)
        cnt = b.bundle.result.cross_count
        cnt = cnt.cpu().numpy() if hasattr(cnt, "cpu") else np.asarray(cnt)
        med = ph.Medium.from_density(n, x, y, freq=136.268e9)
        out = ph.beam_to_receiver(Beam(x0=x_ant, y0=y_ant, kx0=0, ky0=1), med, nf, (-y_ant, "y"), rnum=101, dist=3,
                                  ampbnd=1e-3, rtol=1e-10, atol_r=1e-10, atol_p=1e-10)
        torch.cuda.synchronize()
        ok = (cnt == 1) & (out["status"].cpu().numpy() == ph.STOP_LINE)
        assert ok.sum() > 60                                     # the outer rays are refracted out sideways first
        rk_ph = out["phase"].cpu().numpy()
        worst.append(np.abs(cs[5] - rk_ph)[ok].max())
        assert np.abs(cs[0] - out["x"].cpu().numpy())[ok].max() < 0.1 * (201 / nx)      # crossing points, cm
        rk_phase.append(rk_ph)
        masks.append(ok)
    both = masks[0] & masks[1]
    # the FT-2 profile has a slope discontinuity at the plasma edge, which the two grids sample differently
    print("rk phase, 201 vs 401:", np.abs(rk_phase[0] - rk_phase[1])[both].max(), "parity - rk:", worst)
    assert np.abs(rk_phase[0] - rk_phase[1])[both].max() < 0.15
    assert worst[0] < 2.0 and 1.35 < worst[0] / worst[1] < 2.7      # 0.69 -> 0.44 rad measured (kinked profile: below 2x)


# ---------------------------------------------------------------- X mode with |B| from a tabulated equilibrium
def _eq_dict(eq, freq, r0_cm=55.0, z0_cm=0.0):
    from rayrtrace_b200.physics import ELECTRON_CYCLOTRON_HZ_PER_T
    return dict(psi=eq.psi, fpol=eq.fpol, rleft=eq.rleft, rdim=eq.rdim, zmid=eq.zmid, zdim=eq.zdim, simagx=eq.simagx,
                sibdry=eq.sibdry, r0_cm=r0_cm, z0_cm=z0_cm, y_per_tesla=ELECTRON_CYCLOTRON_HZ_PER_T / freq)


@pytest.mark.parametrize("interp", INTERPS)
def test_equilibrium_table_built_on_the_device_is_bitwise_the_checkers(ph, rk, interp):
    """STB_RK_MODE_X_TABLE: the library evaluates |B|^2 of the tabulated psi(R, Z), F(psi) (the reference's Equilibrium
    fields, equilibrium.py:11-49, bilinear psi) at the medium's nodes on the device; the checker does the same
    operations on the CPU.  Node tables (and their spline coefficients) agree bit for bit, on a non-square grid
    that sticks out of the psi rectangle on one side."""
    freq = 70e9
    eq = ph.Equilibrium.solovev(kappa=1.3, nr=97, nz=81, nf=33)
    eq.fpol = eq.fpol * np.linspace(1.0, 1.02, eq.fpol.size)          # F varies with psi: exercises the F lookup
    xs = np.linspace(-10, 12, 141)                                     # R up to 0.67 m: beyond rleft + rdim = 0.65
    ys = np.linspace(-9, 9, 111)
    X = 0.2 * np.exp(-(xs[None, :] ** 2 + ys[:, None] ** 2) / 20.0)
    med = ph.Medium(X, xs, ys, mode="X", R0=55.0, interp=interp, equilibrium=eq, freq=freq)
    y2 = rk.equilibrium_y2(_eq_dict(eq, freq), xs.size, ys.size, med.xorg, med.yorg, med.dx, med.dy)
    want = rk.prefilter(y2) if interp == "bspline" else y2
    assert med.table("Y2").tobytes() == want.tobytes()
    # the same table handed over as Y2= gives the same medium
    med2 = ph.Medium(X, xs, ys, mode="X", R0=55.0, interp=interp, Y2=y2)
    assert med2.table("Y2").tobytes() == want.tobytes()
    # and it is the Solov'ev field: toroidal part exact, poloidal part to second order in the psi grid spacing
    RR, ZZ = (55.0 + xs[None, :]) * 0.01 + 0 * ys[:, None], ys[:, None] * 0.01 + 0 * xs[None, :]
    inside = (RR < eq.rleft + eq.rdim) & (np.abs(ZZ) < eq.zdim / 2)
    BR, BZ, _ = eq.analytic_B(RR, ZZ)
    F = np.interp(np.clip((eq.analytic["psi_b"] / (eq.analytic["a"] ** 2 * eq.analytic["R0"] ** 2)
                           * (RR ** 2 * ZZ ** 2 / 1.3 ** 2 + (RR ** 2 - 0.55 ** 2) ** 2 / 4)), eq.simagx, eq.sibdry),
                  np.linspace(eq.simagx, eq.sibdry, eq.fpol.size), eq.fpol)
    ypt = ph.ELECTRON_CYCLOTRON_HZ_PER_T / freq
    exact = ypt ** 2 * (BR ** 2 + BZ ** 2 + (F / RR) ** 2)
    assert np.max(np.abs(y2 - exact)[inside] / exact[inside]) < 1e-4


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("absorb", [False, True])
def test_tabulated_field_x_mode_bitwise(ph, rk, interp, absorb):
    xs, X, alpha = _tables(with_alpha=absorb, peak=0.3)
    freq = 60e9
    eq = ph.Equilibrium.solovev(kappa=1.2, psi_b=6e-3, nr=129, nz=129, half=0.12)
    n = 3000
    x0, y0, px, py = _beam(n)
    rng = np.random.default_rng(17)
    xscale = rng.uniform(0.2, 2.4, n)
    yscale = rng.uniform(0.8, 1.2, n)
    amp = rng.uniform(0.5, 1.5, n)
    kw = dict(rtol=1e-8, atol_r=1e-8, atol_p=1e-8, h0=0.05, hmin=1e-10, hmax=0.4, max_steps=5000)
    med = ph.Medium(X, xs, xs, alpha=alpha, mode="X", R0=55.0, interp=interp, equilibrium=eq, freq=freq)
    res = _gpu(ph, med, x0, y0, px, py, xscale=xscale, yscale=yscale, amp=amp, L=48, aux=True,
               nbins=32 if absorb else 0, r_lcs=7.85, rho_max=1.5, **kw)
    y2 = rk.equilibrium_y2(_eq_dict(eq, freq), xs.size, xs.size, med.xorg, med.yorg, med.dx, med.dy)
    om = rk.Medium(X, med.xorg, med.yorg, med.dx, med.dy, alpha=alpha, mode="XT", Y2=y2, interp=INTERP_ID[interp])
    want = rk.trace(om, x0, y0, px, py, xscale=xscale, yscale=yscale, amp=amp, L=48, aux=True,
                    nbins=32 if absorb else 0, r_lcs=7.85, rho_max=1.5, threads=8, **kw)
    _same(res, want)
    assert res.traj.cpu().numpy().tobytes() == want["traj"].tobytes()
    assert want["nacc"].min() > 3 and len(set(want["status"].tolist())) >= 2
    if absorb:
        h = res.hist.cpu().numpy()
        assert np.abs(h - want["hist"]).sum() <= 1e-9 * np.abs(want["hist"]).sum()


def test_tabulated_toroidal_field_reproduces_the_analytic_x_mode(ph):
    """A Y^2 table sampled from the analytic field Y0 R0/(R0 + x) on a fine grid: rays end where the analytic X mode
    puts them (the spline of a smooth table is accurate to ~1e-8 here)."""
    import torch
    xs = np.linspace(-10, 10, 401)
    X = 0.3 * np.exp(-(xs[None, :] ** 2 + xs[:, None] ** 2) / 14.0)
    Y0, R0 = 0.45, 55.0
    Y2 = ((Y0 * R0 / (R0 + xs)) ** 2)[None, :] + 0 * xs[:, None]
    x0, y0, px, py = _beam(2000, seed=3)
    a = ph.Medium(X, xs, xs, mode="X", Y0=Y0, R0=R0)
    b = ph.Medium(X, xs, xs, mode="X", R0=R0, Y2=Y2)
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()
    Na = ph.refractive_index(a, dev(x0), dev(y0))
    Nb = ph.refractive_index(b, dev(x0), dev(y0))
    assert float((Na - Nb).abs().max()) < 1e-9
    ra = ph.trace_rk(a, dev(x0), dev(y0), Na * dev(px), Na * dev(py), rtol=1e-10, stop=(9.0, "y"))
    rb = ph.trace_rk(b, dev(x0), dev(y0), Nb * dev(px), Nb * dev(py), rtol=1e-10, stop=(9.0, "y"))
    torch.cuda.synchronize()
    assert torch.equal(ra.status, rb.status)
    on = ra.status == ph.STOP_LINE                  # the free end state is wherever the last step happened to land
    assert int(on.sum()) > 1000
    assert float((ra.end_state - rb.end_state)[:, on].abs().max()) < 1e-6
    assert float((ra.phi - rb.phi)[on].abs().max()) < 1e-6
