"""Parity of the CUDA path against the oracle, through the C ABI.

Bar: BIT-EXACT on trajectories, step counts, end states, turning indices and
the lax/ph running sums (all IEEE +, *, /, sqrt in the reference's order);
the deposition histogram (extension) within 1e-6 relative L1, the tolerance
BASELINE.json's north_star states (FP64 atomics make its summation order
schedule-dependent).
"""
import numpy as np
import pytest

from _util import KAT_NAMES, ft2_density, random_field, random_launch, traj_digest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def st():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from rayrtrace_b200 import sidtrace
    return sidtrace


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


# ---------------------------------------------------------------- legacy surface
@pytest.mark.parametrize("name", KAT_NAMES)
def test_legacy_trace_kat_bitwise(st, kat, name):
    L, nx, ny, stp, dx, dy, x0, y0, kx0, ky0 = kat[f"{name}_params"]
    L, nx, ny = int(L), int(nx), int(ny)
    x = np.zeros(L); y = np.zeros(L); kx = np.zeros(L); ky = np.zeros(L)
    x[0], y[0], kx[0], ky[0] = x0, y0, kx0, ky0
    ax = np.ascontiguousarray(kat[f"{name}_ax"]); ay = np.ascontiguousarray(kat[f"{name}_ay"])
    n = st.trace(x, y, kx, ky, stp, dx, dy, nx, ny, ax, ay)
    assert isinstance(n, int) and n == int(kat[f"{name}_n"])
    keep = len(kat[f"{name}_x"])
    for got, want in ((x, "x"), (y, "y"), (kx, "kx"), (ky, "ky")):
        assert got[:keep].tobytes() == kat[f"{name}_{want}"].tobytes()
    assert not x[n + 1:].any()          # entries beyond state n stay untouched (sidtracemodule.c:36-48)


def test_legacy_calls_follow_their_field_arrays_and_turning_rays(st, oracle_mod):
    """The per-ray surface caches the field upload per (ax, ay) pair and re-uploads whenever the CONTENT
    changed (exact comparison): new arrays, arrays rescaled in place and single-element edits are all
    seen without any call from the user -- the unmodified ray.py never calls invalidate(); rays that
    wander and turn several times stay bit-exact."""
    rng = np.random.default_rng(21)
    nx, ny, L = 97, 81, 3000

    def both(ax, ay, launch):
        got = [np.zeros(L) for _ in range(4)]
        want = [np.zeros(L) for _ in range(4)]
        for g, w, v in zip(got, want, launch):
            g[0] = w[0] = v
        n = st.trace(*got, 0.35, 1.0, 0.8, nx, ny, ax, ay)
        m = oracle_mod.trace(*want, 0.35, 1.0, 0.8, nx, ny, ax, ay)
        assert n == m
        for g, w in zip(got, want):
            assert g.tobytes() == w.tobytes()
        return n

    ax, ay = random_field(rng, nx, ny, scale=0.05)
    steps = []
    for _ in range(40):       # strong field: most rays wander and turn several times
        x0, y0, kx0, ky0 = (float(v[0]) for v in random_launch(rng, 1, nx, ny, kmag=0.6))
        steps.append(both(ax, ay, (x0, y0, kx0, ky0)))
    assert max(steps) > 300
    launch = (40.3, 11.2, 0.1, 0.9)
    n1 = both(ax, ay, launch)
    bx, by = random_field(rng, nx, ny, scale=0.05)           # new arrays
    n2 = both(bx, by, launch)
    ax *= 1.7; ay *= 0.6                                      # in place, caught by the content sample
    n3 = both(ax, ay, launch)
    ax[int(launch[1]), int(launch[0])] += 0.5                 # one cell, in place: seen as well
    n4 = both(ax, ay, launch)
    ay[int(launch[1]) + 1, int(launch[0])] = -ay[int(launch[1]) + 1, int(launch[0])] - 0.3
    n5 = both(ax, ay, launch)
    st.invalidate()
    assert both(ax, ay, launch) == n5
    assert len({n1, n2, n3, n4, n5}) > 1


def test_legacy_check_and_device_layout(st):
    a = np.arange(24.0).reshape(4, 6)
    assert st.check(6, 4, 2, 3, a) == a[2, 3]
    f = st.Field(a, -a, 7, 5)           # (ny-1, nx-1) = (4, 6)
    assert f.peek(2, 3, 0) == a[2, 3] and f.peek(3, 5, 1) == -a[3, 5]
    f.close()


def test_reference_ray_fixture_through_product_ray(st, ft2_ray):
    """BASELINE config 1: one ray through the FT-2 parabolic density via the
    product's Ray (reference constructor, ray.py:15-111)."""
    from rayrtrace_b200.nfield import Nfield
    from rayrtrace_b200.ray import Ray
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    r = Ray(float(ft2_ray["x0"]), float(ft2_ray["y0"]), 0, 1, float(ft2_ray["freq"]), 1, n_field=nf)
    assert len(r.x) == int(ft2_ray["nsteps"]) == 204
    # the launch refractive index goes through SciPy's spline on both sides: 1e-12, see test_nfield
    np.testing.assert_allclose(r.kx, ft2_ray["ray_kx"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r.x, ft2_ray["ray_x"], rtol=1e-9)
    np.testing.assert_allclose(r.y, ft2_ray["ray_y"], rtol=1e-9)
    np.testing.assert_allclose(r.ph, ft2_ray["ray_ph"], rtol=1e-9)
    assert r.st[-1] == 0 and r.st[0] == float(ft2_ray["st"])


@pytest.mark.parametrize("tag", ["k2", "khalf"])
def test_class_level_step_multiplier(st, ft2_ray, tag):
    """Ray.kStep (ray.py:12, :75: st = n_field.st * Ray.kStep) is honoured by Ray and by trace_rays."""
    from rayrtrace_b200.nfield import Nfield
    from rayrtrace_b200.ray import Ray, trace_rays
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    Ray.kStep = float(ft2_ray[f"{tag}_kstep"])
    try:
        r = Ray(float(ft2_ray["x0"]), float(ft2_ray["y0"]), 0, 1, float(ft2_ray["freq"]), 1, n_field=nf)
        b = trace_rays(nf, [float(ft2_ray["x0"])], [float(ft2_ray["y0"])], [0.0], [1.0], float(ft2_ray["freq"]))
    finally:
        Ray.kStep = 1
    assert len(r.x) == int(ft2_ray[f"{tag}_n"]) == int(b.nsteps[0])
    assert r.st[0] == float(ft2_ray[f"{tag}_st0"])
    np.testing.assert_allclose(r.x, ft2_ray[f"{tag}_x"], rtol=1e-9)
    np.testing.assert_allclose(r.y, ft2_ray[f"{tag}_y"], rtol=1e-9)
    assert r.ph[-1] == pytest.approx(float(ft2_ray[f"{tag}_ph_end"]), rel=1e-9)
    rb = b.ray(0)
    np.testing.assert_allclose(rb.x, ft2_ray[f"{tag}_x"], rtol=1e-9)


def test_reference_ray_fixture_exact_launch_bitwise(st, ft2_ray):
    """Same ray fed the reference's exact launch state: the whole trajectory is bit-identical."""
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    L = 100000                                           # Ray.MaxStep (ray.py:11)
    xs = [np.zeros(L) for _ in range(4)]
    for a, v in zip(xs, ft2_ray["launch"]):
        a[0] = v
    i = st.trace(*xs, float(ft2_ray["st"]), float(ft2_ray["dx0"]), float(ft2_ray["dy0"]), nf.nx, nf.ny, nf.ax, nf.ay)
    assert i == 204
    got = np.stack([a[:i + 1] for a in xs], 1)
    assert got.tobytes() == ft2_ray["grid_traj"].tobytes()


# ---------------------------------------------------------------- batched surface
def _fan_inputs(ft2_fan, tag):
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(ft2_fan[f"{tag}_nx"]), float(ft2_fan[f"{tag}_n0"]))
    nf = Nfield(n, x=x, y=y)
    ins = [np.ascontiguousarray(ft2_fan[f"{tag}_{k}"]) for k in ("xg0", "yg0", "kx0", "ky0")]
    return nf, ins


@pytest.mark.parametrize("tag", ["under", "over"])
@pytest.mark.parametrize("div_ieee", [False, True])
@pytest.mark.parametrize("refill", [True, False])
@pytest.mark.parametrize("device_io", [False, True])
@pytest.mark.parametrize("tiled", [False, True])
def test_fan_fixture_bitwise(st, torch_mod, ft2_fan, tag, div_ieee, refill, device_io, tiled):
    nf, ins = _fan_inputs(ft2_fan, tag)
    L = int(ft2_fan["L"])
    f = nf.device_field()
    if device_io:
        ins = [torch_mod.from_numpy(a).cuda() for a in ins]
    res = st.trace_batch(f, *ins, L=L, st=nf.st, dx=nf.dx[0], dy=nf.dy[0], traj=True,
                         div_ieee=div_ieee, refill=refill, tiled=tiled)
    torch_mod.cuda.synchronize()
    nsteps = res.nsteps.cpu().numpy() if device_io else res.nsteps
    end = res.end_state.cpu().numpy() if device_io else res.end_state
    traj = res.traj_ray_major().cpu().numpy()
    assert res.ray_states(17).cpu().numpy().tobytes() == traj[17, :nsteps[17]].tobytes()
    assert np.array_equal(nsteps, ft2_fan[f"{tag}_nsteps"])
    assert end.tobytes() == ft2_fan[f"{tag}_end"].tobytes()
    assert np.array_equal(traj_digest(traj, nsteps), ft2_fan[f"{tag}_traj_sha"])
    # untouched beyond the last written state
    for r in (0, 17, 95):
        assert not traj[r, nsteps[r] + 1:].any()


@pytest.mark.parametrize("tiled", [False, True])
def test_shaped_density_fan_bitwise(st, torch_mod, rxyten_fan, tiled):
    """The fan through the shaped (shifted / elongated / triangular) density that the reference's own
    rxyten_from_params + on_grid produced (tests/golden/make_golden.py): step counts, end states and
    every trajectory equal to the compiled reference's."""
    from rayrtrace_b200.nfield import Nfield
    x = rxyten_fan["x"]
    nf = Nfield(np.ascontiguousarray(rxyten_fan["density"]), x=x, y=x)
    ins = [np.ascontiguousarray(rxyten_fan[k]) for k in ("xg0", "yg0", "kx0", "ky0")]
    res = st.trace_batch(nf.device_field(), *ins, L=int(rxyten_fan["L"]), st=nf.st, dx=nf.dx[0], dy=nf.dy[0], traj=True,
                         want_extras=True, tiled=tiled)
    torch_mod.cuda.synchronize()
    assert np.array_equal(res.nsteps, rxyten_fan["nsteps"])
    assert res.end_state.tobytes() == rxyten_fan["end"].tobytes()
    assert np.array_equal(traj_digest(res.traj_ray_major().cpu().numpy(), res.nsteps), rxyten_fan["traj_sha"])
    assert int((np.asarray(res.turn_index) >= 0).sum()) >= 16          # the reflected rays have a turning point


def _oracle_vs_gpu(st, oracle, rng, nx, ny, nrays, L, *, stp=0.37, dx=0.8, dy=1.3, st_ray=None, **kw):
    ax, ay = random_field(rng, nx, ny)
    x0, y0, kx0, ky0 = random_launch(rng, nrays, nx, ny)
    if nrays >= 8:   # bad starts and edge starts
        x0[0] = -0.5; y0[1] = np.nan; x0[2] = nx - 1.0; y0[3] = 0.0; x0[4] = np.inf; kx0[5] = np.nan
        x0[6] = nx - 1 - 1e-12; kx0[7] = 0.0
    want = oracle.trace_batch(x0, y0, kx0, ky0, L, stp, dx, dy, nx, ny, ax, ay, st_ray=st_ray)
    f = st.Field(ax, ay, nx, ny)
    res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=stp, dx=dx, dy=dy, st_ray=st_ray, traj=True, **kw)
    traj = res.traj_ray_major().cpu().numpy()
    assert np.array_equal(res.nsteps, want["nsteps"])
    assert res.end_state.tobytes() == want["end"].tobytes()
    assert traj.tobytes() == want["traj"].tobytes()
    f.close()
    return res, want, (ax, ay, x0, y0, kx0, ky0)


@pytest.mark.parametrize("shape", [(2, 2, 1, 4), (5, 3, 33, 16), (64, 80, 1000, 300), (201, 201, 4097, 700),
                                   (33, 17, 31, 1), (33, 17, 32, 2), (1001, 777, 3000, 64)])
@pytest.mark.parametrize("div_ieee", [False, True])
@pytest.mark.parametrize("tiled", [False, True])
def test_random_fields_bitwise_vs_oracle(st, oracle_mod, shape, div_ieee, tiled):
    nx, ny, nrays, L = shape
    rng = np.random.default_rng(nx * 1000 + nrays)
    _oracle_vs_gpu(st, oracle_mod, rng, nx, ny, nrays, L, div_ieee=div_ieee, tiled=tiled)
    _oracle_vs_gpu(st, oracle_mod, rng, nx, ny, nrays, L, div_ieee=div_ieee, tiled=tiled, refill=False)


def test_empty_batch(st):
    f = st.Field(np.zeros((3, 3)), np.zeros((3, 3)))
    z = np.zeros(0)
    res = st.trace_batch(f, z, z, z, z, L=8, st=0.1, dx=1, dy=1)
    assert res.nsteps.shape == (0,) and res.end_state.shape == (4, 0)


def test_awkward_divisors_bitwise(st, oracle_mod):
    """dx, dy whose reciprocals are inexact / near a binade edge: the constant-divisor
    sequence must still equal the IEEE quotient the reference computes."""
    rng = np.random.default_rng(11)
    for dx, dy in ((0.1, 0.3), (np.nextafter(2.0, 0), np.nextafter(1.0, 2)), (1e-3, 7.0), (20.0 / 2000, 20.0 / 1000),
                   (3.0, 1 / 3), (2.0 ** -20, 2.0 ** 20)):
        _oracle_vs_gpu(st, oracle_mod, rng, 40, 30, 500, 120, stp=0.21, dx=dx, dy=dy)
    # divisors outside the guarded range fall back to div.rn.f64
    _oracle_vs_gpu(st, oracle_mod, rng, 40, 30, 100, 50, stp=1e-300, dx=1e-290, dy=1e300)


def test_tiny_and_huge_magnitudes_bitwise(st, oracle_mod):
    """Subnormal / overflowing intermediates take the guarded fallback and still match."""
    rng = np.random.default_rng(5)
    nx, ny, n, L = 20, 20, 256, 40
    ax, ay = random_field(rng, nx, ny, scale=1e-300)
    x0, y0, kx0, ky0 = random_launch(rng, n, nx, ny, kmag=1e-308)
    kx0[::3] = 0.0
    ky0[::5] *= 1e300 * 1e308
    for stp in (1e-3, 1.0, 1e150):
        want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, stp, 0.7, 0.9, nx, ny, ax, ay)
        f = st.Field(ax, ay, nx, ny)
        res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=stp, dx=0.7, dy=0.9, traj=True)
        assert np.array_equal(res.nsteps, want["nsteps"])
        assert res.traj.cpu().numpy().tobytes() == want["traj"].tobytes()
        f.close()


def test_per_ray_step_bitwise(st, oracle_mod):
    """Frequency sweeps give every ray its own st (SURVEY 3.3 iii)."""
    rng = np.random.default_rng(9)
    st_ray = rng.uniform(0.1, 0.6, 700)
    _oracle_vs_gpu(st, oracle_mod, rng, 64, 64, 700, 200, st_ray=st_ray)
    _oracle_vs_gpu(st, oracle_mod, rng, 64, 64, 700, 200, st_ray=st_ray, tiled=True)


def test_heavy_tail_refill_equals_static(st, oracle_mod, torch_mod):
    """A few trapped rays (closed orbits in a focusing well) run to the cap while the
    rest leave early: the persistent refill schedule must give the same bits."""
    nx = ny = 65
    yy, xx = np.meshgrid(np.arange(ny - 1) + 0.5 - 32, np.arange(nx - 1) + 0.5 - 32, indexing="ij")
    ax = np.ascontiguousarray(-0.02 * xx); ay = np.ascontiguousarray(-0.02 * yy)   # harmonic well
    rng = np.random.default_rng(1)
    n = 3000
    x0, y0, kx0, ky0 = random_launch(rng, n, nx, ny, kmag=3.0)
    kx0[::50] *= 0.05; ky0[::50] *= 0.05          # slow rays stay trapped
    L = 2500
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.25, 1.0, 1.0, nx, ny, ax, ay, want_traj=True, nthreads=4)
    assert (want["nsteps"] == L - 1).sum() > 50 and np.median(want["nsteps"]) < L / 4
    f = st.Field(ax, ay, nx, ny)
    for refill in (True, False):
        res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=0.25, dx=1.0, dy=1.0, refill=refill, want_extras=True)
        assert np.array_equal(res.nsteps, want["nsteps"])
        assert res.end_state.tobytes() == want["end"].tobytes()
        for tiled in (False, True):
            res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=0.25, dx=1.0, dy=1.0, refill=refill, traj=True, tiled=tiled)
            assert np.array_equal(res.nsteps, want["nsteps"])
            assert res.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes()
    f.close()


# ---------------------------------------------------------------- fused reductions
def test_extras_bitwise_vs_oracle(st, oracle_mod):
    rng = np.random.default_rng(21)
    nx, ny, n, L = 90, 70, 800, 400
    res, want, _ = _oracle_vs_gpu(st, oracle_mod, rng, nx, ny, n, L, want_extras=True)
    for r in range(n):
        k = int(want["nsteps"][r])
        tp = oracle_mod.turning_point(want["traj"][r], k)
        assert int(res.turn_index[r]) == tp
        if tp >= 0:      # turning point = the reference-identical state at that index, bit for bit
            assert res.turn_state[:, r].tobytes() == want["traj"][r, tp].tobytes()
        else:
            assert np.isnan(res.turn_state[:, r]).all()
        if k >= 1:
            lax, ph = oracle_mod.lax_ph(want["traj"][r, :k], 0.37)
            assert float(res.lax_end[r]).hex() == float(lax[-1]).hex()
            assert float(res.ph_end[r]).hex() == float(ph[-1]).hex()
        else:
            assert res.lax_end[r] == 0.0 and res.ph_end[r] == 0.0


def test_extras_match_reference_ray_properties(st, ft2_ray):
    """lax[-1], ph[-1] of the reference Ray (ray.py:117-126) for BASELINE config 1."""
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    a = [np.array([v]) for v in ft2_ray["launch"]]
    res = st.trace_batch(nf.device_field(), *a, L=1000, st=float(ft2_ray["st"]), dx=float(ft2_ray["dx0"]),
                         dy=float(ft2_ray["dy0"]), want_extras=True)
    assert float(res.lax_end[0]).hex() == float(ft2_ray["ray_lax"][-1]).hex()
    assert float(res.ph_end[0]).hex() == float(ft2_ray["ray_ph"][-1]).hex()


def test_deposition_histogram_vs_oracle(st, oracle_mod, ft2_fan):
    """Extension (no reference code): histogram of amp^2*dl over rho bins, oracle = the same
    definition evaluated on reference-identical trajectories.  Tolerance 1e-6 relative L1
    (north_star); summation order differs (warp aggregation + FP64 atomics)."""
    nf, ins = _fan_inputs(ft2_fan, "over")
    L = int(ft2_fan["L"])
    nb = 48
    rng = np.random.default_rng(2)
    amp = rng.uniform(0.1, 1.0, len(ins[0]))
    want_t = oracle_mod.trace_batch(*ins, L, nf.st, nf.dx[0], nf.dy[0], nf.nx, nf.ny, nf.ax, nf.ay)
    hist = np.zeros(nb)
    for r in range(len(amp)):
        k = int(want_t["nsteps"][r])
        oracle_mod.deposit(want_t["traj"][r, :k], nf.st, amp[r], nf.x[0], nf.y[0], nf.dx[0], nf.dy[0], 7.85, 1.2, nb, hist)
    res = st.trace_batch(nf.device_field(), *ins, L=L, st=nf.st, dx=nf.dx[0], dy=nf.dy[0], nbins=nb, amp=amp,
                         xorg=nf.x[0], yorg=nf.y[0], r_lcs=7.85, rho_max=1.2)
    assert hist.sum() > 0
    assert np.abs(res.hist - hist).sum() / np.abs(hist).sum() < 1e-6
    # accumulates into a caller-provided histogram (not cleared)
    res2 = st.trace_batch(nf.device_field(), *ins, L=L, st=nf.st, dx=nf.dx[0], dy=nf.dy[0], hist=res.hist.copy(), amp=amp,
                          xorg=nf.x[0], yorg=nf.y[0], r_lcs=7.85, rho_max=1.2)
    assert np.abs(res2.hist - 2 * hist).sum() / np.abs(hist).sum() < 2e-6


# ---------------------------------------------------------------- size-independent properties at full size
def test_full_size_invariants(st, oracle_mod, torch_mod):
    """BASELINE config 3 shape (1001^2 grid, 10^6 rays): results are independent of the
    schedule and of how the ray set is sharded; a 2000-ray subsample equals the oracle."""
    from rayrtrace_b200 import synth
    case = synth.beam_case(nx=1001, nrays=1_000_000, seed=0)
    f = case.nfield.device_field()
    dev = [torch_mod.from_numpy(a).cuda() for a in case.launch]
    kw = dict(L=case.L, st=case.st, dx=case.dx, dy=case.dy)
    a = st.trace_batch(f, *dev, **kw)
    b = st.trace_batch(f, *dev, refill=False, div_ieee=True, **kw)
    torch_mod.cuda.synchronize()
    assert torch_mod.equal(a.nsteps, b.nsteps) and torch_mod.equal(a.end_state, b.end_state)
    # shard invariance: 4 contiguous shards traced separately == one launch
    n = case.nrays
    for s in range(4):
        lo, hi = s * n // 4, (s + 1) * n // 4
        part = st.trace_batch(f, *[d[lo:hi].contiguous() for d in dev], **kw)
        assert torch_mod.equal(part.nsteps, a.nsteps[lo:hi])
        assert torch_mod.equal(part.end_state, a.end_state[:, lo:hi])
    # checksum of step counts is the total work
    total = int(a.nsteps.sum().item())
    assert total > 500 * n
    # oracle on a subsample
    idx = np.random.default_rng(0).choice(n, 2000, replace=False)
    sub = [x[idx] for x in case.launch]
    want = oracle_mod.trace_batch(*sub, case.L, case.st, case.dx, case.dy, case.nfield.nx, case.nfield.ny,
                                  case.nfield.ax, case.nfield.ay, want_traj=False, nthreads=4)
    assert np.array_equal(a.nsteps.cpu().numpy()[idx], want["nsteps"])
    assert a.end_state.cpu().numpy()[:, idx].tobytes() == want["end"].tobytes()


def test_vacuum_step_count_property(st):
    """Vacuum: a ray advances k0*st/dx cells per step, so n = ceil(distance / advance)."""
    nx = ny = 257
    f = st.Field(np.zeros((ny - 1, nx - 1)), np.zeros((ny - 1, nx - 1)))
    x0 = np.linspace(5, 250, 4000); y0 = np.zeros(4000)
    adv = 0.4375                              # exactly representable -> no rounding drift
    res = st.trace_batch(f, x0, y0, np.zeros(4000), np.full(4000, adv), L=2000, st=1.0, dx=1.0, dy=1.0)
    assert (res.nsteps == int(np.ceil((ny - 1) / adv))).all()
    assert np.array_equal(res.end_state[0], x0) and (res.end_state[1] == adv * res.nsteps).all()


def test_trace_rays_bundle_matches_per_ray_Ray(st, ft2_ray):
    """The batched launcher (replacement of beam.py:340-342) returns, per ray, exactly what the
    per-ray Ray constructor returns."""
    from rayrtrace_b200.nfield import Nfield
    from rayrtrace_b200.ray import Ray, trace_rays
    n, x, y = ft2_density(int(ft2_ray["nx"]), float(ft2_ray["n0"]))
    nf = Nfield(n, x=x, y=y)
    xs = np.linspace(-5, 5, 9); ys = np.full(9, -9.5)
    ang = np.linspace(-0.2, 0.2, 9)
    freq = float(ft2_ray["freq"])
    b = trace_rays(nf, xs, ys, np.sin(ang), np.cos(ang), freq, amp=np.linspace(0.1, 0.9, 9), want_extras=True)
    assert len(b) == 9
    for r in (0, 4, 8):
        one = Ray(xs[r], ys[r], np.sin(ang[r]), np.cos(ang[r]), freq, 0.5, n_field=nf)
        got = b.ray(r)
        assert len(got.x) == len(one.x) == b.nsteps[r]
        # launch direction normalisation differs by <= 1 ulp (np.linalg.norm vs sqrt), see launch_states
        np.testing.assert_allclose(got.x, one.x, rtol=1e-9)
        np.testing.assert_allclose(got.y, one.y, rtol=1e-9)
        np.testing.assert_allclose(got.ky, one.ky, rtol=1e-9)
        np.testing.assert_allclose(got.ph, one.ph, rtol=1e-9)
        np.testing.assert_allclose(b.result.ph_end[r], one.ph[-1], rtol=1e-9)
        assert got.st[-1] == 0 and got.amp == pytest.approx(0.1 + 0.1 * r)


@pytest.mark.parametrize("with_traj", [False, True])
def test_zero_copy_pinned_host_io_bitwise(st, oracle_mod, with_traj):
    """Tiled schedule + pinned host buffers: the kernel reads launch states and writes per-ray
    results directly in host memory (no staging).  Same bits as the oracle; pageable buffers
    (staged path) give the same bits too."""
    rng = np.random.default_rng(77)
    nx, ny, n, L = 120, 90, 5000 + 13, 260
    ax, ay = random_field(rng, nx, ny)
    launch = list(random_launch(rng, n, nx, ny))
    launch[0][:3] = (-1.0, np.nan, nx + 5.0)
    st_ray = rng.uniform(0.2, 0.5, n)
    want = oracle_mod.trace_batch(*launch, L, 0.3, 0.8, 1.3, nx, ny, ax, ay, st_ray=st_ray, nthreads=4)
    f = st.Field(ax, ay, nx, ny)
    for pin in (True, False):
        ins = [st.pinned_copy(a) if pin else a for a in launch]
        sr = st.pinned_copy(st_ray) if pin else st_ray
        out = st.BatchResult()
        if pin:
            out.nsteps = st.pinned_empty(n, np.int32); out.end_state = st.pinned_empty((4, n))
            out.turn_index = st.pinned_empty(n, np.int32); out.lax_end = st.pinned_empty(n); out.ph_end = st.pinned_empty(n)
            out.nsteps[:] = -7
        res = st.trace_batch(f, *ins, L=L, st=0.3, dx=0.8, dy=1.3, st_ray=sr, tiled=True, want_extras=True,
                             traj=True if with_traj else None, out=out)
        assert np.array_equal(res.nsteps, want["nsteps"])
        assert res.end_state.tobytes() == want["end"].tobytes()
        if with_traj:
            assert res.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes()
        for r in range(0, n, 97):
            k = int(want["nsteps"][r])
            assert int(res.turn_index[r]) == oracle_mod.turning_point(want["traj"][r], k)
            if k:
                lax, ph = oracle_mod.lax_ph(want["traj"][r, :k], st_ray[r])
                assert float(res.ph_end[r]).hex() == float(ph[-1]).hex()
                assert float(res.lax_end[r]).hex() == float(lax[-1]).hex()
    f.close()


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("device_io", [False, True])
def test_receiver_crossing_matches_reference_addpoint(st, torch_mod, ft2_cross, tiled, device_io):
    """Fused receiver-line crossing against what the reference's own Ray.addpoint
    returned (tests/golden/ft2_cross.npz).  Tolerance 1e-9 relative (north_star): the reference
    squares through libm pow there (oracle/crossing.py), so this path is not bitwise."""
    from test_oracle_crossing import cross_case
    nf = cross_case(ft2_cross)
    la = ft2_cross["launch"]
    ins = [np.ascontiguousarray(la[:, k]) for k in range(4)]
    xp, yp = np.ascontiguousarray(ft2_cross["x0"]), np.ascontiguousarray(ft2_cross["y0"])
    if device_io:
        ins = [torch_mod.from_numpy(a).cuda() for a in ins]
        xp, yp = torch_mod.from_numpy(xp).cuda(), torch_mod.from_numpy(yp).cuda()
    res_all = ft2_cross["res"]
    for li, (val, axis) in enumerate(zip(ft2_cross["line_val"], ft2_cross["line_axis"])):
        res = st.trace_batch(nf.device_field(), *ins, L=4096, st=nf.st, dx=nf.dx[0], dy=nf.dy[0], tiled=tiled,
                             traj=True if tiled else None, cross=(float(val), str(axis)), x0_phys=xp, y0_phys=yp)
        get = (lambda a: a.cpu().numpy()) if device_io else (lambda a: a)
        cnt, idx, cs = get(res.cross_count), get(res.cross_index), get(res.cross_state)
        assert np.array_equal(get(res.nsteps), ft2_cross["nkept"])
        want = res_all[li]
        assert np.array_equal(cnt, want[:, 0].astype(int)), (li, cnt, want[:, 0])
        hit = want[:, 0] > 0
        assert np.array_equal(idx[hit], want[hit, 1].astype(int))
        assert (idx[~hit] == -1).all() and np.isnan(cs[:, ~hit]).all()
        np.testing.assert_allclose(cs[:, hit].T, want[hit, 2:8], rtol=1e-9, atol=1e-9)


def test_receiver_crossing_vs_oracle_random(st, oracle_mod):
    """Same, against oracle/crossing.py on a random field (exact hits, both axes, extrapolation)."""
    from oracle.crossing import crossing
    rng = np.random.default_rng(31)
    nx, ny, n, L = 60, 50, 400, 300
    ax, ay = random_field(rng, nx, ny, scale=0.05)
    x0, y0, kx0, ky0 = random_launch(rng, n, nx, ny)
    y0[:50] = 10.0                                   # exact hits of the line y = 2.0 (phys) below
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.4, 0.5, 0.25, nx, ny, ax, ay)
    xp = -3.0 + x0 * 0.5; yp = -0.5 + y0 * 0.25       # physical launch points of an equidistant grid
    f = st.Field(ax, ay, nx, ny)
    for val, axis in ((2.0, "y"), (6.0, "y"), (9.5, "x"), (100.0, "x")):
        res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=0.4, dx=0.5, dy=0.25, cross=(val, axis), x0_phys=xp, y0_phys=yp)
        nhit = 0
        for r in range(n):
            k = int(want["nsteps"][r])
            o = crossing(want["traj"][r, :k], xp[r], yp[r], 0.5, 0.25, 0.4, ax, ay, val, axis)
            assert int(res.cross_count[r]) == o["count"], (val, axis, r)
            if o["count"]:
                nhit += 1
                assert int(res.cross_index[r]) == o["index"]
                assert int(res.cross_flags[r]) == o["exact"] + 2 * o["extrapolated"]
                np.testing.assert_allclose(res.cross_state[:, r], [o[q] for q in ("x", "y", "kx", "ky", "lax", "ph")],
                                           rtol=1e-9, atol=1e-9)
        assert nhit > 20 or val == 100.0
    f.close()


def test_beam_field_result_matches_reference(st, beam_fix):
    """the product's Beam (one batched launch instead of the serial Ray loop,
    beam.py:340-342) followed by field_r / field_result reproduces the reference's interferometer
    signal for a TX/RX antenna pair on the FT-2 profile.  Tolerance 1e-9 relative: the launch
    refractive index and addpoint's squares are tolerance-level paths (see test_nfield, oracle/crossing.py)."""
    from scipy.interpolate import interp1d
    from rayrtrace_b200.beam import Beam, field_result
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    x0, y0 = float(beam_fix["x0"]), float(beam_fix["y0"])
    tx = Beam(x0=x0, y0=y0, kx0=0, ky0=1)
    rx = Beam(x0=x0, y0=-y0, kx0=0, ky0=-1)
    wf = interp1d(beam_fix["wf_x"], beam_fix["wf_y"], kind="cubic")
    out = tx.rays(rnum=int(beam_fix["rnum"]), wave_front=wf, dist=3, ampbnd=1e-3, n_field=nf)
    np.testing.assert_allclose(np.stack(out[:6]), beam_fix["launch"], rtol=1e-13, atol=1e-13)
    assert np.array_equal(tx.bundle.nsteps, beam_fix["nkept"])
    # the fused device crossing agrees with the host addpoint on the receiver line
    b2 = Beam(x0=x0, y0=y0, kx0=0, ky0=1)
    b2.rays(rnum=int(beam_fix["rnum"]), wave_front=wf, dist=3, ampbnd=1e-3, n_field=nf, receiver=(-y0, "y"))
    field, xo, x_max, ph_max, signal, props = field_result(tx, rx)
    hit = b2.bundle.result.cross_count == 1
    np.testing.assert_allclose(np.sort(b2.bundle.result.cross_state[0][hit]), xo, rtol=1e-9)
    np.testing.assert_allclose(xo, beam_fix["xout"], rtol=1e-9)
    np.testing.assert_allclose(field, beam_fix["field"], rtol=1e-7, atol=1e-9)
    assert abs(signal - complex(beam_fix["signal"])) < 1e-7 * abs(complex(beam_fix["signal"]))
    assert float(ph_max) == pytest.approx(float(beam_fix["ph_max"]), rel=1e-9)
    assert x_max == pytest.approx(float(beam_fix["x_max"]), rel=1e-8, abs=1e-9)


def test_config2_thousand_ray_beam_deposition(st, oracle_mod):
    """BASELINE config 2: a 10^3-ray Gaussian beam launched by Beam.rays (restated, beam.py:278-344)
    from an FT-2 antenna; trajectories bit-identical to the oracle and the deposition profile
    (extension) within 1e-6 L1 of the oracle's."""
    from rayrtrace_b200.beam import Beam
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(201, 4e13)
    nf = Nfield(n, x=x, y=y)
    xa = 2.3
    b = Beam(x0=xa, y0=-np.sqrt(100 - xa ** 2), kx0=0, ky0=1)
    b.rays(rnum=1000, dist=3, ampbnd=1e-3, n_field=nf, tiled=False)
    bun = b.bundle
    assert len(bun) == 1000 and bun.nsteps.min() > 100 and np.median(bun.nsteps) > 180
    L = bun.result.traj.shape[1]
    want = oracle_mod.trace_batch(bun.xg0, bun.yg0, bun.kx0, bun.ky0, L, bun.st, nf.dx[0], nf.dy[0], nf.nx, nf.ny, nf.ax, nf.ay,
                                  nthreads=4)
    assert np.array_equal(bun.nsteps, want["nsteps"])
    assert bun.result.traj.cpu().numpy().tobytes() == want["traj"].tobytes()
    nb = 40
    hist = np.zeros(nb)
    for r in range(1000):
        oracle_mod.deposit(want["traj"][r, :want["nsteps"][r]], bun.st, bun.amp[r], nf.x[0], nf.y[0], nf.dx[0], nf.dy[0],
                           7.85, 1.3, nb, hist)
    res = st.trace_batch(nf.device_field(), bun.xg0, bun.yg0, bun.kx0, bun.ky0, L=L, st=bun.st, dx=nf.dx[0], dy=nf.dy[0],
                         nbins=nb, amp=bun.amp, xorg=nf.x[0], yorg=nf.y[0], r_lcs=7.85, rho_max=1.3)
    assert np.abs(res.hist - hist).sum() / hist.sum() < 1e-6


def test_config4_sweep_invariants(st, oracle_mod, torch_mod):
    """BASELINE config 4: launch-angle x frequency sweep over an over-dense (reflecting) profile,
    10^7 rays with per-ray step (frequency) -- a cutoff-heavy, ragged workload (launch states
    prepared on the device).  Schedules agree bit for bit; a 3000-ray subsample equals the oracle;
    a third of the rays reflect."""
    from rayrtrace_b200 import synth
    case = synth.sweep_case(nx=201, nrays=10_000_000, seed=4, device_launch=True)
    rng = np.random.default_rng(4)
    fscale = rng.uniform(100e9, 160e9, case.nrays) / case.freq          # per-ray frequency, 100-160 GHz
    st_ray = case.st / fscale                                           # st ~ 1/w (den.py:107); k0 ~ w
    launch = [case.launch[0], case.launch[1], case.launch[2] * fscale, case.launch[3] * fscale]
    f = case.nfield.device_field()
    dev = [torch_mod.from_numpy(np.ascontiguousarray(a)).cuda() for a in launch]
    dst = torch_mod.from_numpy(st_ray).cuda()
    kw = dict(L=case.L, st=case.st, dx=case.dx, dy=case.dy, st_ray=dst, want_extras=True)
    a = st.trace_batch(f, *dev, **kw)
    b = st.trace_batch(f, *dev, refill=True, **kw)
    c = st.trace_batch(f, *dev, tiled=True, div_ieee=True, **kw)
    torch_mod.cuda.synchronize()
    for o in (b, c):
        assert torch_mod.equal(a.nsteps, o.nsteps) and torch_mod.equal(a.end_state, o.end_state)
        assert torch_mod.equal(a.ph_end, o.ph_end) and torch_mod.equal(a.turn_index, o.turn_index)
    refl = float((a.end_state[3] < 0).double().mean().item())
    assert 0.2 < refl < 0.6
    idx = rng.choice(case.nrays, 3000, replace=False)
    want = oracle_mod.trace_batch(*[x[idx] for x in launch], case.L, case.st, case.dx, case.dy, case.nfield.nx, case.nfield.ny,
                                  case.nfield.ax, case.nfield.ay, st_ray=st_ray[idx], nthreads=4)
    assert np.array_equal(a.nsteps.cpu().numpy()[idx], want["nsteps"])
    assert a.end_state.cpu().numpy()[:, idx].tobytes() == want["end"].tobytes()
    ph = a.ph_end.cpu().numpy()[idx]
    for j in range(0, 3000, 150):
        k = int(want["nsteps"][j])
        assert float(ph[j]).hex() == float(oracle_mod.lax_ph(want["traj"][j, :k], st_ray[idx][j])[1][-1]).hex()


@pytest.mark.parametrize("tile", [32, 64, 128])
@pytest.mark.parametrize("refill", [None, True, False])
def test_tile_sizes_bitwise(st, oracle_mod, tile, refill):
    """Tiled trajectory layout with 64- and 128-ray tiles ([ray/T][step][ray%T][4]): same bits."""
    rng = np.random.default_rng(tile)
    nx, ny, n, L = 70, 60, 1000 + 37, 220
    ax, ay = random_field(rng, nx, ny)
    launch = random_launch(rng, n, nx, ny)
    want = oracle_mod.trace_batch(*launch, L, 0.37, 0.8, 1.3, nx, ny, ax, ay)
    f = st.Field(ax, ay, nx, ny)
    res = st.trace_batch(f, *launch, L=L, st=0.37, dx=0.8, dy=1.3, traj=True, tiled=True, tile=tile, refill=refill)
    assert tuple(res.traj.shape) == ((n + tile - 1) // tile, L, tile, 4)
    assert np.array_equal(res.nsteps, want["nsteps"])
    assert res.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes()
    r = n - 3
    assert res.ray_states(r).cpu().numpy().tobytes() == want["traj"][r, :want["nsteps"][r]].tobytes()
    f.close()


@pytest.mark.parametrize("tag", ["tiltp", "tiltm"])
def test_tilted_beam_matches_reference(st, beam_fix, tag):
    """Launch-angle axis of BASELINE config 4: a beam tilted with the Beam.k0 setter (beam.py:96-100):
    launch set, kept lengths and the interferometer signal against the reference's."""
    from rayrtrace_b200.beam import Beam, field_result
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    x0, y0, th = float(beam_fix["x0"]), float(beam_fix["y0"]), float(beam_fix[f"{tag}_angle"])
    for device in (False, True):
        tx = Beam(x0=x0, y0=y0, kx0=0, ky0=1)
        tx.k0 = (np.sin(th), np.cos(th))
        rx = Beam(x0=x0, y0=-y0, kx0=0, ky0=-1)
        out = tx.rays(rnum=61, dist=3, ampbnd=1e-3, n_field=nf, receiver=(-y0, "y") if device else None)
        np.testing.assert_allclose(np.stack(out[:6]), beam_fix[f"{tag}_launch"], rtol=1e-12, atol=1e-12)
        assert np.array_equal(tx.bundle.nsteps, beam_fix[f"{tag}_nkept"])
        field, xo, x_max, ph_max, signal, _ = field_result(tx, rx, device=device)
        want = complex(beam_fix[f"{tag}_signal"])
        assert abs(signal - want) < 1e-6 * abs(want), device
        assert float(ph_max) == pytest.approx(float(beam_fix[f"{tag}_ph_max"]), rel=1e-8)
        assert x_max == pytest.approx(float(beam_fix[f"{tag}_x_max"]), rel=1e-6, abs=1e-8)


@pytest.mark.parametrize("tag", ["f100", "f160"])
def test_other_probing_frequencies_match_reference(st, beam_fix, tag):
    """Frequency axis of BASELINE config 4: field (den.py:84-107), step and beam at 100 and 160 GHz."""
    from rayrtrace_b200.beam import Beam, field_result
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    f = float(beam_fix[f"{tag}_freq"])
    nf = Nfield(n, x=x, y=y, freq=f)
    assert float(nf.st) == float(beam_fix[f"{tag}_st"])
    x0, y0 = float(beam_fix["x0"]), float(beam_fix["y0"])
    for device in (False, True):
        tx = Beam(x0=x0, y0=y0, kx0=0, ky0=1, freq=f)
        rx = Beam(x0=x0, y0=-y0, kx0=0, ky0=-1, freq=f)
        out = tx.rays(rnum=61, dist=3, ampbnd=1e-3, n_field=nf, receiver=(-y0, "y") if device else None)
        np.testing.assert_allclose(np.stack(out[:6]), beam_fix[f"{tag}_launch"], rtol=1e-12, atol=1e-12)
        assert np.array_equal(tx.bundle.nsteps, beam_fix[f"{tag}_nkept"])
        _, _, x_max, ph_max, signal, _ = field_result(tx, rx, device=device)
        want = complex(beam_fix[f"{tag}_signal"])
        assert abs(signal - want) < 1e-6 * abs(want), device
        assert float(ph_max) == pytest.approx(float(beam_fix[f"{tag}_ph_max"]), rel=1e-8)


def test_ft2_phases_driver_matches_reference(st, ft2_phases_fix):
    """Top of the reference's call stack (ft2.phases, ft2.py:25-71): density scan x seven TX/RX
    antenna pairs -> interferometer phases, every ray traced on the GPU through the batched Beam.
    The fixture is the reference's own ft2.phases output (tests/golden/make_golden.py)."""
    from rayrtrace_b200 import ft2
    n, x, y = ft2_density(int(ft2_phases_fix["nx"]), float(ft2_phases_fix["n0"]))
    # GPU tail (everything O(rays) on the device) / the two launches + NumPy tail / materialised Ray objects
    for device in (True, "host-tail", False):
        ph_max, ph_sig = ft2.phases(n, x, y, a_set=ft2_phases_fix["a_set"], nray=int(ft2_phases_fix["nray"]), device=device)
        assert ph_max.shape == ph_sig.shape == (3, 7)
        d = np.angle(np.exp(1j * (ph_sig - ft2_phases_fix["ph_by_sig"])))      # compare angles modulo 2 pi
        assert np.abs(d).max() < 1e-6
        np.testing.assert_allclose(ph_max, ft2_phases_fix["ph_by_max"], rtol=1e-6)


def test_ft2_phases_driver_matches_reference_in_dense_plasmas(st, ft2_phases_fix):
    """The same driver at peak densities of 0.52 and 0.78 of critical: strongly bent, partly folded
    beams (several 'pages' in field_r, rays missing the receiver)."""
    from rayrtrace_b200 import ft2
    n, x, y = ft2_density(int(ft2_phases_fix["nx"]), float(ft2_phases_fix["n0"]))
    for device in (True, "host-tail", False):
        ph_max, ph_sig = ft2.phases(n, x, y, a_set=ft2_phases_fix["dense_a_set"], nray=int(ft2_phases_fix["nray"]), device=device)
        d = np.angle(np.exp(1j * (ph_sig - ft2_phases_fix["dense_ph_by_sig"])))
        assert np.abs(d).max() < 1e-5, (device, d)
        np.testing.assert_allclose(ph_max, ft2_phases_fix["dense_ph_by_max"], rtol=1e-6)


def test_cap_at_reference_maxstep(st, oracle_mod):
    """A ray that never leaves the grid stops at the cap exactly like the reference at
    Ray.MaxStep = 100000 (ray.py:11, sidtracemodule.c:36): i = L-1 = 99999, state L-1 written."""
    nx = ny = 9
    ax = np.zeros((ny - 1, nx - 1)); ay = np.zeros((ny - 1, nx - 1))
    L = 100000
    xs = [np.zeros(L) for _ in range(4)]
    xs[0][0], xs[1][0] = 4.5, 4.5                      # k = 0: stands still forever
    n = st.trace(*xs, 0.1, 1.0, 1.0, nx, ny, ax, ay)
    assert n == L - 1 and xs[0][L - 1] == 4.5 and xs[1][L - 1] == 4.5
    # batch: one standing ray among moving ones, both schedules, cap reported per ray
    x0 = np.array([4.5, 1.0, 2.0, 7.9]); y0 = np.array([4.5, 1.0, 2.0, 0.1])
    kx0 = np.array([0.0, 1.0, -1.0, 0.5]); ky0 = np.array([0.0, 0.5, 2.0, -1.0])
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.1, 1.0, 1.0, nx, ny, ax, ay, want_traj=False)
    f = st.Field(ax, ay, nx, ny)
    for refill in (True, False):
        res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=0.1, dx=1.0, dy=1.0, refill=refill)
        assert np.array_equal(res.nsteps, want["nsteps"]) and res.nsteps[0] == L - 1
        assert res.end_state.tobytes() == want["end"].tobytes()
    f.close()


def test_equal_phase_lookup_matches_reference_interp1d(st, ft2_cross):
    """LOOKUP pass (second launch of Beam.field_r's device path): each ray, with its crossing point
    inserted as after the reference's addpoint, sampled at two target phases == the reference's
    interp1d(r.ph, r.x / r.y, fill_value='extrapolate') (beam.py:410-413).  Tolerance 1e-9."""
    from test_oracle_crossing import cross_case
    nf = cross_case(ft2_cross)
    la = ft2_cross["launch"]
    ins = [np.ascontiguousarray(la[:, k]) for k in range(4)]
    xp, yp = np.ascontiguousarray(ft2_cross["x0"]), np.ascontiguousarray(ft2_cross["y0"])
    checked = 0
    for li, (val, axis) in enumerate(zip(ft2_cross["line_val"], ft2_cross["line_axis"])):
        t = ft2_cross["look_t"][li]                       # [nray, 2], NaN where the reference had no single crossing
        ok = ~np.isnan(t[:, 0])
        targets = np.ascontiguousarray(np.where(ok[:, None], t, 0.0).T)
        for refill in (True, False):
            res = st.trace_batch(nf.device_field(), *ins, L=4096, st=nf.st, dx=nf.dx[0], dy=nf.dy[0],
                                 cross=(float(val), str(axis)), x0_phys=xp, y0_phys=yp, lookup_ph=targets, refill=refill)
            assert np.array_equal(res.cross_count, ft2_cross["res"][li][:, 0].astype(int))
            np.testing.assert_allclose(res.lookup_xy[:, ok].T, ft2_cross["look_xy"][li][ok], rtol=1e-9, atol=1e-9)
        checked += int(ok.sum())
    assert checked > 60


def test_field_r_device_matches_reference(st, beam_fix):
    """Beam.field_r without Ray objects (fused crossing + equal-phase lookup launch) reproduces the
    reference's field on the receiver line and the interferometer signal (beam.py:347-472)."""
    from scipy.interpolate import interp1d
    from rayrtrace_b200.beam import Beam, field_r_device, field_result
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(int(beam_fix["nx"]), float(beam_fix["n0"]))
    nf = Nfield(n, x=x, y=y)
    x0, y0 = float(beam_fix["x0"]), float(beam_fix["y0"])
    wf = interp1d(beam_fix["wf_x"], beam_fix["wf_y"], kind="cubic")
    tx = Beam(x0=x0, y0=y0, kx0=0, ky0=1)
    rx = Beam(x0=x0, y0=-y0, kx0=0, ky0=-1)
    tx.rays(rnum=int(beam_fix["rnum"]), wave_front=wf, dist=3, ampbnd=1e-3, n_field=nf, keep_traj=False, receiver=(-y0, "y"))
    field, xo, amp0, amp1, ph, pnt = field_r_device(tx, -y0)
    np.testing.assert_allclose(xo, beam_fix["xout"], rtol=1e-9)
    f2, x2, x_max, ph_max, signal, props = field_result(tx, rx, device=True)
    np.testing.assert_allclose(f2, beam_fix["field"], rtol=1e-7, atol=1e-12)      # TX field x RX pattern (beam.py:451)
    assert abs(signal - complex(beam_fix["signal"])) < 1e-7 * abs(complex(beam_fix["signal"]))
    assert float(ph_max) == pytest.approx(float(beam_fix["ph_max"]), rel=1e-9)
    # the same with every O(rays) array and reduction on the GPU (receiver_field_device, stb_tube_field,
    # stb_hole_fill, stb_overlap, Levenberg-Marquardt over stb_gauss_fit_sums instead of curve_fit)
    f3, x3, x_max3, ph_max3, signal3, props3 = field_result(tx, rx, device="gpu")
    np.testing.assert_allclose(x3, beam_fix["xout"], rtol=1e-9)
    np.testing.assert_allclose(f3, beam_fix["field"], rtol=1e-7, atol=1e-12)
    assert abs(signal3 - complex(beam_fix["signal"])) < 1e-7 * abs(complex(beam_fix["signal"]))
    assert x_max3 == pytest.approx(float(beam_fix["x_max"]), rel=1e-8, abs=1e-9)
    np.testing.assert_allclose(props3, beam_fix["props"], rtol=1e-7)           # MINPACK stops at its own tolerance
    assert float(ph_max3) == pytest.approx(float(beam_fix["ph_max"]), rel=1e-9)
    # and equals the host path (Ray objects + addpoint + interp1d) on the same beam
    tx2 = Beam(x0=x0, y0=y0, kx0=0, ky0=1)
    tx2.rays(rnum=int(beam_fix["rnum"]), wave_front=wf, dist=3, ampbnd=1e-3, n_field=nf)
    fh = tx2.field_r(-y0)
    np.testing.assert_allclose(field, fh[0], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(amp1, fh[3], rtol=1e-7)


@pytest.mark.parametrize("refill", [True, False])
@pytest.mark.parametrize("tiled", [False, True])
def test_no_out_of_bounds_writes_canaries(st, torch_mod, oracle_mod, refill, tiled):
    """Own bounds check (compute-sanitizer is closed on this pool): every output buffer sits inside
    a larger arena filled with a canary; after the launch the canaries are intact and the payload
    equals the oracle.  Ray count deliberately not a multiple of 32 / 128."""
    rng = np.random.default_rng(123)
    nx, ny, n, L = 50, 40, 1000 + 77, 90
    ax, ay = random_field(rng, nx, ny)
    launch = random_launch(rng, n, nx, ny)
    want = oracle_mod.trace_batch(*launch, L, 0.37, 0.8, 1.3, nx, ny, ax, ay)
    f = st.Field(ax, ay, nx, ny)
    dev = [torch_mod.from_numpy(a).cuda() for a in launch]
    pad = 4096
    CAN = -7.25e300

    def arena(numel, dtype):
        buf = torch_mod.full((numel + 2 * pad,), CAN if dtype == torch_mod.float64 else -77777, dtype=dtype, device="cuda")
        return buf, buf[pad:pad + numel]

    T = 128
    nt = (n + T - 1) // T
    tshape = (nt, L, T, 4) if tiled else (n, L, 4)
    bufs = {k: arena(int(np.prod(shp)), dt) for k, (shp, dt) in {
        "traj": (tshape, torch_mod.float64), "nsteps": ((n,), torch_mod.int32), "end": ((4, n), torch_mod.float64),
        "turn": ((n,), torch_mod.int32), "lax": ((n,), torch_mod.float64), "ph": ((n,), torch_mod.float64)}.items()}
    out = st.BatchResult(nsteps=bufs["nsteps"][1], end_state=bufs["end"][1].view(4, n), turn_index=bufs["turn"][1],
                         lax_end=bufs["lax"][1], ph_end=bufs["ph"][1])
    res = st.trace_batch(f, *dev, L=L, st=0.37, dx=0.8, dy=1.3, traj=bufs["traj"][1].view(*tshape), tiled=tiled, tile=T,
                         refill=refill, want_extras=True, out=out)
    torch_mod.cuda.synchronize()
    for k, (buf, _) in bufs.items():
        can = CAN if buf.dtype == torch_mod.float64 else -77777
        assert bool((buf[:pad] == can).all()) and bool((buf[-pad:] == can).all()), f"canary of {k} overwritten"
    assert np.array_equal(res.nsteps.cpu().numpy(), want["nsteps"])
    assert res.end_state.cpu().numpy().tobytes() == want["end"].tobytes()
    traj = res.traj_ray_major().cpu().numpy()
    for r in range(n):
        k = int(want["nsteps"][r])
        assert traj[r, :k + 1].tobytes() == want["traj"][r, :k + 1].tobytes()
        assert (traj[r, k + 1:] == CAN).all()          # never written beyond state i
    if tiled:                                          # slots of the partial last tile stay untouched
        assert bool((res.traj[-1, :, n % T:, :] == CAN).all())
    f.close()


def test_plain_c_caller_of_the_abi(st, tmp_path):
    """The boundary is a C ABI: a C program (no Python, no torch) binds include/sidtrace_b200.h,
    traces KAT G1 through stb_trace and stb_trace_batch and compares with the published bits."""
    import os
    import subprocess
    from rayrtrace_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "abi_kat"
    lib = _cabi.LIB_PATH
    subprocess.check_call([os.environ.get("CC", "gcc"), "-O1", "-std=c11", os.path.join(root, "tests", "c", "abi_kat.c"),
                           lib, "-lm", "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout


def test_guard_special_values_bitwise(st, oracle_mod):
    """The loop guard `0 <= x < nx-1 and 0 <= y < ny-1` (sidtracemodule.c:36) is evaluated with
    integer compares on the device: -0.0 (passes: -0.0 >= 0), tiny negatives in (-1, 0), +-inf,
    NaN payloads, subnormals, values at and next to the upper bound, and values >= 2^31 must
    all decide exactly as the FP64 compares of the reference do."""
    nx, ny, L = 12, 9, 40
    rng = np.random.default_rng(8)
    ax, ay = random_field(rng, nx, ny, scale=0.05)
    ax[0, :] = -0.0; ay[0, :] = -0.0                     # a -0.0 field row keeps -0.0 states alive
    nan2 = np.frombuffer(np.uint64(0x7ff0000000000001).tobytes(), np.float64)[0]     # NaN with hi word 0x7ff00000
    nan3 = np.frombuffer(np.uint64(0xfff8000000000000).tobytes(), np.float64)[0]
    specials = [-0.0, 0.0, 5e-324, -5e-324, -1e-300, -0.5, -0.999999, -1.0, np.nextafter(nx - 1.0, 0), nx - 1.0,
                np.nextafter(nx - 1.0, 99), 2.0 ** 31, 2.0 ** 31 - 0.5, 2.0 ** 40, 1e300, np.inf, -np.inf, np.nan, nan2, nan3,
                np.nextafter(ny - 1.0, 0), ny - 1.0]
    x0, y0, kx0, ky0 = [], [], [], []
    for v in specials:
        for w in (0.25, -0.0):
            x0 += [v, w]; y0 += [w, v]
            kx0 += [-0.0, 0.3]; ky0 += [0.2, -0.0]
    x0, y0, kx0, ky0 = (np.array(a, np.float64) for a in (x0, y0, kx0, ky0))
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.5, 1.0, 1.0, nx, ny, ax, ay)
    assert 0 < (want["nsteps"] > 0).sum() < len(x0)
    f = st.Field(ax, ay, nx, ny)
    for kw in (dict(), dict(refill=True), dict(tiled=True), dict(div_ieee=True), dict(want_extras=True)):
        res = st.trace_batch(f, x0, y0, kx0, ky0, L=L, st=0.5, dx=1.0, dy=1.0, traj=True, **kw)
        assert np.array_equal(res.nsteps, want["nsteps"]), kw
        assert res.end_state.tobytes() == want["end"].tobytes(), kw
        assert res.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes(), kw
    f.close()


def test_trajectory_export_to_host_packed(st, oracle_mod):
    """export.trace_to_host: chunked trace + pack + device-to-host copy on a second stream; the
    packed states of every ray equal the oracle's kept trajectory bit for bit."""
    from rayrtrace_b200 import export
    rng = np.random.default_rng(55)
    nx, ny, n, L = 80, 64, 5000 + 9, 200
    ax, ay = random_field(rng, nx, ny)
    launch = random_launch(rng, n, nx, ny)
    launch[0][:2] = (-3.0, np.nan)                                  # rays with zero kept states
    want = oracle_mod.trace_batch(*launch, L, 0.37, 0.8, 1.3, nx, ny, ax, ay, nthreads=4)
    f = st.Field(ax, ay, nx, ny)
    for chunk in (640, 100000):                                      # many chunks (ragged last one) / a single chunk
        off, states, nsteps, end = export.trace_to_host(f, *launch, L=L, st=0.37, dx=0.8, dy=1.3, chunk_rays=chunk)
        assert np.array_equal(nsteps, want["nsteps"]) and end.tobytes() == want["end"].tobytes()
        assert off[-1] == want["nsteps"].sum() == states.shape[0]
        for r in range(0, n, 7):
            k = int(want["nsteps"][r])
            assert states[off[r]:off[r + 1]].tobytes() == want["traj"][r, :k].tobytes()
    f.close()


def test_device_launch_states_match_host(st, torch_mod, ft2_ray):
    """stb_launch_states (ray.py:34-68 for a whole beam, on the GPU) against the host restatement:
    grid coordinates bit-identical (scipy interp1d arithmetic, incl. extrapolation outside the grid),
    wave vectors to 1e-13 (same spline coefficients, FITPACK evaluation restated on the device)."""
    from rayrtrace_b200.nfield import Nfield
    from rayrtrace_b200.ray import launch_states, launch_states_device
    rng = np.random.default_rng(17)
    for nx, n0 in ((201, 4e13), (129, 3.0e14)):
        n, x, y = ft2_density(nx, n0)
        nf = Nfield(n, x=x, y=y)
        m = 20000
        x0 = rng.uniform(-10.5, 10.5, m); y0 = rng.uniform(-10.5, 10.5, m)     # some points outside the grid
        x0[:4] = (x[0], x[-1], x[5], -10.0); y0[:4] = (y[0], y[-1], y[7], 10.0)  # exactly on nodes / edges
        ang = rng.uniform(0, 2 * np.pi, m)
        kx0 = np.cos(ang) * rng.uniform(0.1, 3, m); ky0 = np.sin(ang) * rng.uniform(0.1, 3, m)
        inside = (np.abs(x0) < 7.5) & (np.abs(y0) < 7.5) if n0 > 2e14 else np.ones(m, bool)   # avoid n > n_c (NaN index)
        freq = float(ft2_ray["freq"])
        hx, hy, hkx, hky, _ = launch_states(nf, x0, y0, kx0, ky0, freq)
        dx_, dy_, dkx, dky, _ = launch_states_device(nf, x0, y0, kx0, ky0, freq)
        torch_mod.cuda.synchronize()
        assert dx_.cpu().numpy().tobytes() == hx.tobytes() and dy_.cpu().numpy().tobytes() == hy.tobytes()
        ok = ~np.isnan(hkx)
        assert ok.mean() > 0.5
        np.testing.assert_allclose(dkx.cpu().numpy()[ok], hkx[ok], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dky.cpu().numpy()[ok], hky[ok], rtol=1e-12, atol=1e-12)
        assert np.array_equal(np.isnan(dkx.cpu().numpy()), np.isnan(hkx)) or (~ok).sum() < 5


def test_fuzz_random_configurations_bitwise(st, oracle_mod, torch_mod):
    """80 random configurations (grid shape, ray count, cap, step, cell sizes, per-ray steps, schedule,
    layout, tile, division mode, host/device/pinned I/O, fused outputs): everything bit-identical
    to the oracle."""
    import os
    rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "2026")))      # FUZZ_ITERS / FUZZ_SEED: long one-off runs
    for it in range(int(os.environ.get("FUZZ_ITERS", "80"))):
        nx, ny = int(rng.integers(2, 90)), int(rng.integers(2, 90))
        n = int(rng.choice([1, 2, 31, 32, 33, 127, 129, 255, 257, 700, 2049]))
        L = int(rng.choice([1, 2, 3, 17, 64, 300]))
        stp = float(rng.uniform(0.05, 1.5)); dx = float(rng.uniform(0.1, 3.0)); dy = float(rng.uniform(0.1, 3.0))
        ax, ay = random_field(rng, nx, ny, scale=float(rng.choice([0.0, 0.01, 0.5])))
        launch = random_launch(rng, n, nx, ny, kmag=float(rng.uniform(0.2, 3)))
        if n > 4:
            launch[0][rng.integers(0, n)] = rng.choice([-0.0, -1e-9, np.nan, nx - 1.0, 1e20])
        st_ray = rng.uniform(0.05, 1.0, n) if rng.random() < 0.3 else None
        tiled = bool(rng.random() < 0.5)
        kw = dict(refill=[None, True, False][int(rng.integers(0, 3))], div_ieee=bool(rng.random() < 0.3),
                  tiled=tiled, tile=int(rng.choice([32, 64, 128, 256])), want_extras=bool(rng.random() < 0.5))
        mode = int(rng.integers(0, 3))          # 0 host pageable, 1 host pinned, 2 device
        want = oracle_mod.trace_batch(*launch, L, stp, dx, dy, nx, ny, ax, ay, st_ray=st_ray)
        f = st.Field(ax, ay, nx, ny)
        ins, sr = list(launch), st_ray
        if mode == 1:
            ins = [st.pinned_copy(a) for a in launch]
            sr = st.pinned_copy(st_ray) if st_ray is not None else None
        elif mode == 2:
            ins = [torch_mod.from_numpy(a).cuda() for a in launch]
            sr = torch_mod.from_numpy(st_ray).cuda() if st_ray is not None else None
        res = st.trace_batch(f, *ins, L=L, st=stp, dx=dx, dy=dy, st_ray=sr, traj=True, **kw)
        get = (lambda a: a.cpu().numpy()) if mode == 2 else (lambda a: np.asarray(a))
        ctx = (it, nx, ny, n, L, kw, mode)
        assert np.array_equal(get(res.nsteps), want["nsteps"]), ctx
        assert get(res.end_state).tobytes() == want["end"].tobytes(), ctx
        assert res.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes(), ctx
        if kw["want_extras"]:
            ti = get(res.turn_index)
            r = int(rng.integers(0, n))
            k = int(want["nsteps"][r])
            assert int(ti[r]) == oracle_mod.turning_point(want["traj"][r], k), ctx
            if k >= 1:
                s_r = stp if st_ray is None else st_ray[r]
                assert float(get(res.ph_end)[r]).hex() == float(oracle_mod.lax_ph(want["traj"][r, :k], s_r)[1][-1]).hex(), ctx
        f.close()


@pytest.mark.parametrize("rnum", [7, 251, 1000, 70001])
def test_beam_launch_conditions_device_match_host(st, beam_fix, rnum):
    """stb_beam_launch (Beam.rays launch conditions, beam.py:293-331, on the GPU) against the host
    restatement, which is itself pinned bit for bit to the reference's Beam.rays: 1e-11."""
    from scipy.interpolate import interp1d
    from rayrtrace_b200.beam import Beam
    b = Beam(x0=float(beam_fix["x0"]), y0=float(beam_fix["y0"]), kx0=0.05, ky0=1)
    b.k0 = (0.05, 1.0)                                  # slightly tilted beam: exercises the frame rotation
    wf = interp1d(beam_fix["wf_x"], beam_fix["wf_y"], kind="cubic")
    host = b.launch_conditions(rnum, wf, 0.998)
    devo = [t.cpu().numpy() for t in b.launch_conditions_device(rnum, wf, 0.998)]
    for h, d, name in zip(host[:5], devo, ("x", "y", "kx", "ky", "amp")):
        np.testing.assert_allclose(d, h, rtol=1e-11, atol=1e-11, err_msg=name)


# ---------------------------------------------------------------- round 2: hoisted table, exact arenas, the reference's cap
@pytest.mark.parametrize("mode", ["endpoints", "traj", "traj-tiled", "extras", "cross", "hist"])
@pytest.mark.parametrize("refill", [True, False])
def test_hoisted_table_equals_plain_cells(st, torch_mod, oracle_mod, mode, refill):
    """Launches with one global step read the four cell-only products of sidtracemodule.c:41-45 from a
    per-(field, st) table (HOIST kernels); the same launch with hoist=False recomputes them per step.
    Same operations on the same operands: every output is bit-identical (the histogram to 1e-12: atomics)."""
    rng = np.random.default_rng(31)
    nx, ny, n, L = 150, 131, 3001, 500
    ax, ay = random_field(rng, nx, ny, scale=0.02)
    x0, y0, kx0, ky0 = random_launch(rng, n, nx, ny)
    x0[0] = -1.0; y0[1] = np.nan; kx0[2] = 0.0
    f = st.Field(ax, ay, nx, ny)
    kw = dict(L=L, st=0.41, dx=0.9, dy=1.1, refill=refill)
    if mode == "traj":
        kw.update(traj=True)
    elif mode == "traj-tiled":
        kw.update(traj=True, tiled=True)
    elif mode == "extras":
        kw.update(want_extras=True)
    elif mode == "cross":
        kw.update(cross=(70.0, "y"), x0_phys=x0 * 0.9 - 3.0, y0_phys=y0 * 1.1 + 2.0)
    elif mode == "hist":
        kw.update(nbins=32, amp=rng.uniform(0.5, 1, n), xorg=-60.0, yorg=-70.0, r_lcs=50.0, rho_max=1.4)
    a = st.trace_batch(f, x0, y0, kx0, ky0, hoist=True, **kw)
    b = st.trace_batch(f, x0, y0, kx0, ky0, hoist=False, **kw)
    torch_mod.cuda.synchronize()
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.41, 0.9, 1.1, nx, ny, ax, ay)
    assert np.array_equal(a.nsteps, want["nsteps"]) and np.array_equal(b.nsteps, want["nsteps"])
    assert a.end_state.tobytes() == want["end"].tobytes() == b.end_state.tobytes()
    if mode.startswith("traj"):
        assert a.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes()
    for name in ("turn_index", "turn_state", "lax_end", "ph_end", "cross_count", "cross_index", "cross_flags", "cross_state"):
        va, vb = getattr(a, name), getattr(b, name)
        if va is not None:
            assert np.asarray(va).tobytes() == np.asarray(vb).tobytes(), name
    if mode == "hist":
        assert np.abs(a.hist - b.hist).sum() <= 1e-12 * np.abs(b.hist).sum()
    f.close()


@pytest.mark.parametrize("mode", ["endpoints", "traj", "traj-tiled", "extras", "cross", "lookup", "hist"])
@pytest.mark.parametrize("refill", [True, False])
@pytest.mark.parametrize("hoist", [True, False])
def test_ymajor_cells_equal_the_reference_layout(st, torch_mod, oracle_mod, mode, refill, hoist):
    """cells="y" gathers from the y-major copy of the cell table (STB_F_YMAJOR: cell (iy, ix) at ix*(ny-1) + iy)
    instead of the reference's a[iy*(nx-1) + ix] (sidtracemodule.c:38).  Same cells: every output is
    bit-identical to the x-major launch and to the oracle, on a non-square grid, hoisted or not."""
    rng = np.random.default_rng(77)
    nx, ny, n, L = 97, 161, 3003, 500
    ax, ay = random_field(rng, nx, ny, scale=0.02)
    x0, y0, kx0, ky0 = random_launch(rng, n, nx, ny)
    x0[0] = -1.0; y0[1] = np.nan; kx0[2] = 0.0; x0[3] = -0.0; y0[4] = -0.0
    f = st.Field(ax, ay, nx, ny)
    kw = dict(L=L, st=0.41, dx=0.9, dy=1.1, refill=refill, hoist=hoist)
    xp, yp = x0 * 0.9 - 3.0, y0 * 1.1 + 2.0
    if mode == "traj":
        kw.update(traj=True)
    elif mode == "traj-tiled":
        kw.update(traj=True, tiled=True)
    elif mode == "extras":
        kw.update(want_extras=True)
    elif mode in ("cross", "lookup"):
        kw.update(cross=(70.0, "y"), x0_phys=xp, y0_phys=yp)
        if mode == "lookup":
            c0 = st.trace_batch(f, x0, y0, kx0, ky0, **kw)
            ph = np.nan_to_num(c0.cross_state[5], nan=50.0)
            kw.update(lookup_ph=np.ascontiguousarray(np.stack([0.6 * ph, 1.2 * ph + 1.0])))
    elif mode == "hist":
        kw.update(nbins=32, amp=rng.uniform(0.5, 1, n), xorg=-60.0, yorg=-70.0, r_lcs=50.0, rho_max=1.4)
    a = st.trace_batch(f, x0, y0, kx0, ky0, cells="y", **kw)
    b = st.trace_batch(f, x0, y0, kx0, ky0, cells="x", **kw)
    torch_mod.cuda.synchronize()
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.41, 0.9, 1.1, nx, ny, ax, ay)
    assert np.array_equal(a.nsteps, want["nsteps"]) and np.array_equal(b.nsteps, want["nsteps"])
    assert a.end_state.tobytes() == want["end"].tobytes() == b.end_state.tobytes()
    if mode.startswith("traj"):
        assert a.traj_ray_major().cpu().numpy().tobytes() == want["traj"].tobytes()
    for name in ("turn_index", "turn_state", "lax_end", "ph_end", "cross_count", "cross_index", "cross_flags", "cross_state",
                 "lookup_xy"):
        va, vb = getattr(a, name), getattr(b, name)
        if va is not None:
            assert np.asarray(va).tobytes() == np.asarray(vb).tobytes(), name
    if mode == "hist":
        assert np.abs(a.hist - b.hist).sum() <= 1e-12 * np.abs(b.hist).sum()
    assert st.pick_cells(kx0, ky0) in ("x", "y")
    up = np.abs(ky0) + 1.0
    assert st.pick_cells(kx0 * 0.1, up) == "y" and st.pick_cells(up, ky0 * 0.1) == "x"
    assert st.pick_cells(torch_mod.from_numpy(kx0 * 0.1).cuda(), torch_mod.from_numpy(up).cuda()) == "y"
    f.close()


def test_hoisted_tables_of_many_steps_share_one_field(st, torch_mod, oracle_mod):
    """The field keeps four hoisted tables (LRU): six step values traced round-robin, on two streams,
    evict and rebuild them; every launch matches the oracle."""
    rng = np.random.default_rng(8)
    nx, ny, n, L = 90, 70, 2000, 260
    ax, ay = random_field(rng, nx, ny, scale=0.03)
    launch = random_launch(rng, n, nx, ny)
    d_in = [torch_mod.from_numpy(a).cuda() for a in launch]
    f = st.Field(ax, ay, nx, ny)
    steps = [0.2, 0.25, 0.3, 0.35, 0.4, 0.45]
    want = {s: oracle_mod.trace_batch(*launch, L, s, 1.0, 1.0, nx, ny, ax, ay) for s in steps}
    streams = [torch_mod.cuda.Stream(), torch_mod.cuda.Stream()]
    got = []
    for rep in range(3):
        for k, s in enumerate(steps):
            with torch_mod.cuda.stream(streams[k % 2]):
                got.append((s, st.trace_batch(f, *d_in, L=L, st=s, dx=1.0, dy=1.0)))
    torch_mod.cuda.synchronize()
    for s, r in got:
        assert np.array_equal(r.nsteps.cpu().numpy(), want[s]["nsteps"])
        assert r.end_state.cpu().numpy().tobytes() == want[s]["end"].tobytes()
    f.close()


def test_queue_heads_are_not_shared_by_launches_in_flight(st, torch_mod, oracle_mod):
    """More persistent launches in flight than the field has queue-head slots (64), on several
    streams: a slot is reused only after the launch that used it last has finished, so no ray is
    traced twice or skipped (the histogram would show it)."""
    rng = np.random.default_rng(77)
    nx, ny, n, L = 64, 64, 20000, 400
    ax, ay = random_field(rng, nx, ny, scale=0.01)
    launch = random_launch(rng, n, nx, ny)
    d_in = [torch_mod.from_numpy(a).cuda() for a in launch]
    f = st.Field(ax, ay, nx, ny)
    want = oracle_mod.trace_batch(*launch, L, 0.3, 1.0, 1.0, nx, ny, ax, ay)
    streams = [torch_mod.cuda.Stream() for _ in range(4)]
    outs = []
    for k in range(200):
        with torch_mod.cuda.stream(streams[k % 4]):
            outs.append(st.trace_batch(f, *d_in, L=L, st=0.3, dx=1.0, dy=1.0, refill=True, nbins=8, xorg=-32.0, yorg=-32.0,
                                       r_lcs=30.0, rho_max=1.5))
    torch_mod.cuda.synchronize()
    ref_hist = outs[0].hist.cpu().numpy()
    for r in outs:
        assert np.array_equal(r.nsteps.cpu().numpy(), want["nsteps"])
        assert np.abs(r.hist.cpu().numpy() - ref_hist).sum() <= 1e-12 * np.abs(ref_hist).sum()
    f.close()


@pytest.mark.parametrize("tile", [32, 256])
@pytest.mark.parametrize("nrays", [1, 255, 1000])
def test_exact_arena_heavy_tailed_bitwise(st, torch_mod, oracle_mod, tile, nrays):
    """trace_exact: a first pass yields the step counts, stb_traj_plan sorts the rays by length into
    tiles sized for their longest ray, a second pass writes every kept state into the packed arena.
    Heavy-tailed lengths (strong random field, a cap that some rays hit, bad starts): every ray's states
    0..i equal the oracle's, the arena holds exactly sum(rows_t) * T states, and the plan is a permutation."""
    rng = np.random.default_rng(1000 + nrays)
    nx, ny, L = 97, 81, 3000
    ax, ay = random_field(rng, nx, ny, scale=0.05)
    launch = list(random_launch(rng, nrays, nx, ny, kmag=0.6))
    if nrays > 8:
        launch[0][3] = -5.0; launch[1][5] = np.nan
    d_in = [torch_mod.from_numpy(a).cuda() for a in launch]
    f = st.Field(ax, ay, nx, ny)
    res = st.trace_exact(f, *d_in, L=L, st=0.35, dx=1.0, dy=0.8, tile=tile, want_extras=True)
    torch_mod.cuda.synchronize()
    want = oracle_mod.trace_batch(*launch, L, 0.35, 1.0, 0.8, nx, ny, ax, ay)
    ns = res.nsteps.cpu().numpy()
    assert np.array_equal(ns, want["nsteps"])
    assert res.end_state.cpu().numpy().tobytes() == want["end"].tobytes()
    plan = res.plan
    perm = plan.perm.cpu().numpy()
    assert sorted(perm[perm >= 0].tolist()) == list(range(nrays))
    assert np.all(np.diff(ns[perm[perm >= 0]]) <= 0)                        # longest first
    toff = plan.tile_off.cpu().numpy()
    rows = np.diff(toff)
    assert plan.total_rows == toff[-1] == rows.sum()
    for t in range(len(rows)):
        members = perm[t * tile:(t + 1) * tile]
        assert rows[t] == ns[members[members >= 0]].max() + 1
    assert tuple(res.traj.shape) == (plan.total_rows, tile, 4)
    assert plan.total_rows * tile <= (ns + 1).sum() + tile * L and plan.arena_bytes == res.traj.numel() * 8
    step = max(1, nrays // 97)
    for r in list(range(0, nrays, step)) + [int(np.argmax(ns)), int(np.argmin(ns))]:
        got = res.ray_states(r, int(ns[r]) + 1).cpu().numpy()
        assert got.tobytes() == want["traj"][r, :ns[r] + 1].tobytes(), r
    f.close()


def test_exact_arena_export_packs_kept_states(st, torch_mod, oracle_mod):
    """stb_pack_traj_plan: the packed arena -> [sum(nsteps), 4] in launch order, the kept states the
    reference's Ray holds (ray.py:99-102)."""
    import ctypes
    rng = np.random.default_rng(4)
    nx, ny, n, L = 60, 50, 700, 900
    ax, ay = random_field(rng, nx, ny, scale=0.04)
    launch = random_launch(rng, n, nx, ny, kmag=0.7)
    d_in = [torch_mod.from_numpy(a).cuda() for a in launch]
    f = st.Field(ax, ay, nx, ny)
    res = st.trace_exact(f, *d_in, L=L, st=0.3, dx=1.0, dy=1.0, tile=64)
    want = oracle_mod.trace_batch(*launch, L, 0.3, 1.0, 1.0, nx, ny, ax, ay)
    ns = res.nsteps.to(torch_mod.int64)
    off = torch_mod.zeros(n + 1, dtype=torch_mod.int64, device=ns.device)
    off[1:] = torch_mod.cumsum(ns, 0)
    out = torch_mod.empty((int(off[-1]), 4), dtype=torch_mod.float64, device=ns.device)
    lib = st._cabi.lib
    st._cabi.check(lib.stb_pack_traj_plan(res.traj.data_ptr(), 64, res.plan.slot_of.data_ptr(), res.plan.tile_off.data_ptr(), n,
                                          res.nsteps.data_ptr(), off.data_ptr(), out.data_ptr(),
                                          ctypes.c_void_p(torch_mod.cuda.current_stream().cuda_stream)))
    torch_mod.cuda.synchronize()
    exp = np.concatenate([want["traj"][r, :want["nsteps"][r]] for r in range(n)])
    assert out.cpu().numpy().tobytes() == exp.tobytes()
    f.close()


@pytest.mark.parametrize("tag", ["near", "ring"])
@pytest.mark.parametrize("tiled", [True, False])
def test_beam_rays_run_to_the_reference_cap(st, long_rays_fix, tag, tiled):
    """Beam.rays with NO max_steps equals the reference for rays far beyond 2*nx steps (the reference's
    cap is Ray.MaxStep = 100000, ray.py:11): a near-cutoff fan whose central rays crawl for 500+ steps
    and a cavity in which every ray bounces for ~2000 steps on a 201^2 grid.  Fixture: the reference's
    own Beam.rays / Ray (tests/golden/make_golden.py make_long_rays)."""
    from rayrtrace_b200.beam import Beam
    fx = long_rays_fix
    nf = _long_rays_field(fx, tag)
    assert float(nf.st) == float(fx[f"{tag}_st"])
    xa, ya = fx[f"{tag}_ant"]
    b = Beam(x0=float(xa), y0=float(ya), kx0=0, ky0=1)
    out = b.rays(rnum=int(fx[f"{tag}_rnum"]), dist=3, ampbnd=float(fx[f"{tag}_ampbnd"]), n_field=nf, tiled=tiled)
    np.testing.assert_allclose(np.stack(out[:5]), fx[f"{tag}_launch"], rtol=1e-12, atol=1e-13)
    bun = b.bundle
    nk = fx[f"{tag}_nkept"]
    assert nk.max() > 511                                   # beyond the old default cap
    assert np.array_equal(bun.nsteps, nk)
    np.testing.assert_allclose(bun.result.ph_end, fx[f"{tag}_ph_end"], rtol=1e-9)
    np.testing.assert_allclose(bun.result.lax_end, fx[f"{tag}_lax_end"], rtol=1e-9)
    for r in (0, len(nk) // 2, int(np.argmax(nk))):
        ray = bun.ray(r)
        assert len(ray.x) == nk[r]
        np.testing.assert_allclose([ray.x[-1], ray.y[-1], ray.kx[-1], ray.ky[-1]], fx[f"{tag}_last"][r], rtol=1e-8, atol=1e-9)


def _long_rays_field(fx, tag):
    from rayrtrace_b200.nfield import Nfield
    nx = int(fx["nx"])
    if tag == "near":
        n, x, y = ft2_density(nx, float(fx["near_n0"]))
    else:
        x = np.linspace(-10.0, 10.0, nx)
        X, Y = np.meshgrid(x, x)
        n, y = 4.0e14 * np.exp(-((np.sqrt(X ** 2 + Y ** 2) - 6.0) / 2.5) ** 2) + 1e12, x
    return Nfield(n, x=x, y=y)


@pytest.mark.parametrize("tag", ["near", "ring"])
def test_long_reference_rays_exact_launch_bitwise(st, torch_mod, long_rays_fix, tag):
    """The reference's exact launch state (its own k0 after the refractive-index scaling) through
    trace_exact with L = Ray.MaxStep: every kept array of every ray -- 500 to 2400 points on a 201^2
    grid -- hashes to the reference Ray's (physical x, y as ray.py:105-106 forms them)."""
    from rayrtrace_b200.ray import Ray
    fx = long_rays_fix
    nf = _long_rays_field(fx, tag)
    lx, ly = fx[f"{tag}_launch"][0], fx[f"{tag}_launch"][1]
    xg = np.ascontiguousarray(nf.x2grid(lx), np.float64)
    yg = np.ascontiguousarray(nf.y2grid(ly), np.float64)
    k0 = fx[f"{tag}_k0"]
    d_in = [torch_mod.from_numpy(np.ascontiguousarray(a)).cuda() for a in (xg, yg, k0[:, 0], k0[:, 1])]
    res = st.trace_exact(nf.device_field(), *d_in, L=int(Ray.MaxStep), st=float(nf.st), dx=float(nf.dx[0]), dy=float(nf.dy[0]))
    torch_mod.cuda.synchronize()
    ns = res.nsteps.cpu().numpy()
    assert np.array_equal(ns, fx[f"{tag}_nkept"])
    import hashlib
    for r in range(len(ns)):
        t = res.ray_states(r).cpu().numpy()
        x = lx[r] + (t[:, 0] - t[0, 0]) * nf.dx[0]
        y = ly[r] + (t[:, 1] - t[0, 1]) * nf.dy[0]
        got = hashlib.sha256(np.ascontiguousarray(np.stack([x, y, t[:, 2], t[:, 3]], 1)).tobytes()).hexdigest()
        assert got == str(fx[f"{tag}_sha"][r]), r


def test_plain_c_two_rank_caller_of_the_comm_abi(st, torch_mod, tmp_path):
    """The multi-GPU entry points are a C ABI too (include/sidtrace_b200_comm.h): tests/c/comm_2rank.c forks
    two processes, one GPU each -- NCCL from a unique id over a pipe, block-interleaved shards of one beam,
    all-reduced histogram, summaries gathered and unshuffled on rank 0 -- and compares with one GPU tracing
    the whole beam.  Needs two GPUs (gpurun --gpus 2); skipped otherwise."""
    import os
    import subprocess
    from rayrtrace_b200 import _cabi
    if torch_mod.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "comm_2rank"
    lib = _cabi.LIB_PATH
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call([os.environ.get("CC", "gcc"), "-O1", os.path.join(root, "tests", "c", "comm_2rank.c"),
                           "-I" + os.path.join(cuda, "include"), lib, "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                           "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-identical" in r.stdout and "MISMATCH" not in r.stdout


def test_gpu_tail_equals_host_tail_on_a_folded_many_page_beam(st, torch_mod):
    """The device-side O(rays) tail against the NumPy / SciPy one on a beam that needs it all: 20 001 rays
    through a plasma at 0.8 of critical (hundreds of pages, rays that miss the receiver, holes)."""
    from rayrtrace_b200.beam import Beam, field_r_device, field_result, receiver_field_device
    from rayrtrace_b200.nfield import Nfield
    n, x, y = ft2_density(201, 1.85e14)
    nf = Nfield(n, x=x, y=y)
    xa = -2.3
    ya = -np.sqrt(100 - xa ** 2)
    tx = Beam(x0=xa, y0=ya, kx0=0, ky0=1)
    rx = Beam(x0=xa, y0=-ya, kx0=0, ky0=-1)
    tx.rays(rnum=20001, n_field=nf, keep_traj=False, receiver=(float(-ya), "y"))
    fh, xh, a0h, a1h, phh, pnth = field_r_device(tx, -ya)
    d = receiver_field_device(tx, -ya)
    g = {k: v.cpu().numpy() for k, v in d.items()}
    assert len(xh) == len(g["xout"]) > 15000
    np.testing.assert_allclose(g["xout"], xh, rtol=0, atol=0)
    np.testing.assert_allclose(g["ph"], phh, rtol=0, atol=0)
    assert np.array_equal(g["pnt"], pnth)
    np.testing.assert_allclose(g["amp1"], a1h, rtol=1e-9)
    np.testing.assert_allclose(g["field"][:, 0] + 1j * g["field"][:, 1], fh, rtol=1e-9, atol=1e-13)
    host = field_result(tx, rx, device=True)
    gpu = field_result(tx, rx, device="gpu")
    np.testing.assert_allclose(gpu[0], host[0], rtol=1e-9, atol=1e-13)
    assert gpu[2] == pytest.approx(host[2], rel=1e-10)
    assert abs(gpu[4] - host[4]) < 1e-9 * abs(host[4])
    if host[5] is not None:
        np.testing.assert_allclose(gpu[5], host[5], rtol=1e-6)
        assert float(gpu[3]) == pytest.approx(float(host[3]), rel=1e-8)


@pytest.mark.parametrize("world,rows,elem,n_global,stride_pad", [(8, 4, 8, 256 * 37 + 101, 0), (8, 1, 4, 256 * 37 + 101, 0), (3, 4, 8, 256 * 10, 3),
                                                                  (2, 1, 4, 256 * 5 + 1, 1), (5, 2, 8, 7, 0)])
def test_unshuffle_blocks_on_one_gpu(st, torch_mod, world, rows, elem, n_global, stride_pad):
    """stb_unshuffle_blocks (gathered shards -> global ray order) needs no communicator: gathered buffers are made on
    one GPU and the result compared with the index map written out in NumPy -- whole runs (the 16-byte vector
    kernel), an odd destination stride (the element-wise fallback), a partial last block, a chunk offset g0."""
    import ctypes
    from rayrtrace_b200 import _cabi
    B = 256
    nb = -(-n_global // B)
    lblocks = -(-nb // world)
    dt, tdt = (np.float64, torch_mod.float64) if elem == 8 else (np.int32, torch_mod.int32)
    rng = np.random.default_rng(world * 100 + rows)
    src = rng.integers(-2 ** 30, 2 ** 30, (world, rows, lblocks * B)).astype(dt)
    for chunk_lo, chunk_n in ((0, lblocks), (1, max(lblocks - 1, 0))):         # the whole shard; a chunk that starts at local block 1
        if chunk_n == 0:
            continue
        stride = n_global + stride_pad
        dst = torch_mod.full((rows, stride), -7, dtype=tdt, device="cuda")
        part = np.ascontiguousarray(src[:, :, chunk_lo * B:(chunk_lo + chunk_n) * B])
        d_src = torch_mod.from_numpy(part).cuda()
        g0 = chunk_lo * world * B
        _cabi.check(_cabi.lib.stb_unshuffle_blocks(d_src.data_ptr(), dst.data_ptr(), world, rows, chunk_n, B, elem, g0, n_global,
                                                   stride, ctypes.c_void_p(torch_mod.cuda.current_stream().cuda_stream)))
        torch_mod.cuda.synchronize()
        want = np.full((rows, stride), -7, dtype=dt)
        for r in range(world):
            for lb in range(chunk_n):
                g = g0 + (lb * world + r) * B
                m = min(B, n_global - g)
                if m > 0:
                    want[:, g:g + m] = part[r, :, lb * B:lb * B + m]
        assert np.array_equal(dst.cpu().numpy(), want)


@pytest.mark.parametrize("mode", ["endpoints", "extras", "cross", "hist", "traj", "traj-tiled"])
@pytest.mark.parametrize("device_io,refill", [(False, None), (True, None), (True, True), (True, False)])
def test_sort_rays_returns_the_callers_order(st, torch_mod, oracle_mod, mode, device_io, refill):
    """trace_batch(sort_rays=True) traces a random-order launch set along a Morton curve of launch cells (coherent
    gathers) and hands every per-ray result back in the caller's order; trajectories stay in traced order and are
    read through slot_of.  Bit-identical to the unsorted launch and to the oracle."""
    rng = np.random.default_rng(41)
    nx, ny, n, L = 150, 131, 3001, 400
    ax, ay = random_field(rng, nx, ny, scale=0.02)
    x0, y0, kx0, ky0 = random_launch(rng, n, nx, ny)
    x0[0] = -1.0; y0[1] = np.nan; kx0[2] = 0.0
    f = st.Field(ax, ay, nx, ny)
    up = (lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()) if device_io else (lambda a: a)
    kw = dict(L=L, st=0.41, dx=0.9, dy=1.1)
    if mode == "traj":
        kw.update(traj=True)
    elif mode == "traj-tiled":
        kw.update(traj=True, tiled=True)
    elif mode == "extras":
        kw.update(want_extras=True)
    elif mode == "cross":
        kw.update(cross=(70.0, "y"), x0_phys=up(x0 * 0.9 - 3.0), y0_phys=up(y0 * 1.1 + 2.0))
    elif mode == "hist":
        kw.update(nbins=32, amp=up(rng.uniform(0.5, 1, n)), xorg=-60.0, yorg=-70.0, r_lcs=50.0, rho_max=1.4)
    ins = [up(a) for a in (x0, y0, kx0, ky0)]
    if not device_io:
        with pytest.raises(ValueError):                 # the order is applied inside the kernel: device I/O only
            st.trace_batch(f, *ins, sort_rays=True, **kw)
        f.close()
        return
    a = st.trace_batch(f, *ins, sort_rays=True, refill=refill, **kw)
    b = st.trace_batch(f, *ins, refill=refill, **kw)
    torch_mod.cuda.synchronize()
    host = lambda v: v.cpu().numpy() if device_io else np.asarray(v)
    want = oracle_mod.trace_batch(x0, y0, kx0, ky0, L, 0.41, 0.9, 1.1, nx, ny, ax, ay)
    assert np.array_equal(host(a.nsteps), want["nsteps"])
    assert host(a.end_state).tobytes() == want["end"].tobytes()
    for name in ("turn_index", "turn_state", "lax_end", "ph_end", "cross_count", "cross_index", "cross_flags", "cross_state"):
        va, vb = getattr(a, name), getattr(b, name)
        if va is not None:
            assert host(va).tobytes() == host(vb).tobytes(), name
    if mode == "hist":
        assert np.abs(host(a.hist) - host(b.hist)).sum() <= 1e-12 * np.abs(host(b.hist)).sum()
    if mode.startswith("traj"):
        if mode == "traj-tiled":
            assert sorted(host(a.order).tolist()) == list(range(n)) and a.slot_of is not None
        else:
            assert a.slot_of is None                     # ray-major rows sit at the ray's own index
        for r in (0, 1, 7, n // 2, n - 1):
            k = int(want["nsteps"][r])
            assert host(a.ray_states(r, k + 1)).tobytes() == want["traj"][r, :k + 1].tobytes()
        assert host(a.traj_ray_major()).tobytes() == want["traj"].tobytes()
    # a precomputed order can be reused
    order = st.coherent_order(*ins)
    c = st.trace_batch(f, *ins, order=order, refill=refill, **{k: v for k, v in kw.items() if k != "traj"})
    assert np.array_equal(host(c.nsteps), want["nsteps"])
    f.close()
