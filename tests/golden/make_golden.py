"""Generates the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden.py

Everything is produced by reference code:
  * kat.npz       -- G1..G6 of SURVEY.md section 4.1: the compiled reference
                     extension (sidtracemodule.c) called exactly as ray.py:80-83
                     calls it.
  * ft2_ray.npz   -- BASELINE config 1: one ray through the FT-2 parabolic
                     density, via the reference's den.Nfield and ray.Ray
                     (both traceType 'c' and the pure-Python twin ray.py:85-96).
  * ft2_fan.npz   -- a 96-ray fan through an under-dense and an over-dense
                     (reflecting) FT-2 profile: reference Nfield + compiled trace.
Densities are analytic (parameters stored), so fixtures stay small; the
reference-derived field is pinned by its step `st`, SHA-256 digests of ax/ay
and a strided sample.
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle import refenv  # noqa: E402

warnings.simplefilter("ignore")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ft2_density(nx, n0, edge=1e12, r_dia=7.85, half=10.0):
    """n = n0 * max(0, 1 - r^2/r_dia^2) + edge on x, y = linspace(-half, half, nx)  """
    x = np.linspace(-half, half, nx)
    y = np.linspace(-half, half, nx)
    X, Y = np.meshgrid(x, y)
    n = n0 * np.maximum(0, 1 - (X ** 2 + Y ** 2) / r_dia ** 2) + edge
    return n, x, y


def run_ref_trace(ref, L, x0, y0, kx0, ky0, st, dx, dy, nx, ny, ax, ay):
    x = np.zeros(L); y = np.zeros(L); kx = np.zeros(L); ky = np.zeros(L)
    x[0], y[0], kx[0], ky[0] = x0, y0, kx0, ky0
    n = ref.trace(x, y, kx, ky, st, dx, dy, nx, ny, ax, ay)
    return n, x, y, kx, ky


def make_kat(ref):
    out = {}
    cases = {
        # name: (L, nx, ny, ax, ay, st, dx, dy, x0, y0, kx0, ky0)
        "G1": (4096, 11, 9, np.full((8, 10), 0.5), np.full((8, 10), -0.25), 0.5, 1.0, 2.0, 1.2, 3.7, 0.3, 0.9),
        "G2": (4096, 33, 17, np.zeros((16, 32)), np.zeros((16, 32)), 0.5, 1.0, 1.0, 4.0, 0.0, 0.25, 1.0),
        "G3": (4096, 21, 41, np.zeros((40, 20)), np.full((40, 20), -0.8), 0.125, 0.5, 0.5, 10.25, 3.3, 0.0625, 2.0),
        "G4": (4096, 7, 5, np.tile(np.arange(1.0, 7.0), (4, 1)), -np.tile(np.arange(1.0, 5.0)[:, None], (1, 6)),
               0.25, 1.0, 1.0, 0.5, 0.5, 1.0, 2.0),
        "G5": (8, 101, 101, np.zeros((100, 100)), np.zeros((100, 100)), 0.5, 1.0, 1.0, 50.0, 50.0, 0.0, 1.0),
        "G6a": (4096, 101, 101, np.zeros((100, 100)), np.zeros((100, 100)), 0.5, 1.0, 1.0, -0.5, 50.0, 0.0, 1.0),
        "G6b": (4096, 101, 101, np.zeros((100, 100)), np.zeros((100, 100)), 0.5, 1.0, 1.0, np.nan, 50.0, 0.0, 1.0),
        "G6c": (4096, 101, 101, np.zeros((100, 100)), np.zeros((100, 100)), 0.5, 1.0, 1.0, 100.0, 50.0, 0.0, 1.0),
    }
    for name, (L, nx, ny, ax, ay, st, dx, dy, x0, y0, kx0, ky0) in cases.items():
        ax = np.ascontiguousarray(ax, np.float64)
        ay = np.ascontiguousarray(ay, np.float64)
        n, x, y, kx, ky = run_ref_trace(ref, L, x0, y0, kx0, ky0, st, dx, dy, nx, ny, ax, ay)
        keep = min(L, n + 2)   # states 0..n plus one untouched entry
        out[f"{name}_params"] = np.array([L, nx, ny, st, dx, dy, x0, y0, kx0, ky0], np.float64)
        out[f"{name}_ax"] = ax
        out[f"{name}_ay"] = ay
        out[f"{name}_n"] = np.int64(n)
        out[f"{name}_x"] = x[:keep]
        out[f"{name}_y"] = y[:keep]
        out[f"{name}_kx"] = kx[:keep]
        out[f"{name}_ky"] = ky[:keep]
        print(name, "n =", n)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)


def make_ft2_ray(den, ray):
    nx = 201
    n, x, y = ft2_density(nx, 4e13)
    nf = den.Nfield(n, x=x, y=y)                       # reference field producer, default freq 136.26 GHz
    freq = 136.268e9                                   # Beam's launch frequency (beam.py:30): mismatch replicated
    ray.Ray.traceType = "c"
    rc = ray.Ray(-3.0, -9.5, 0, 1, freq, 1, n_field=nf)
    ray.Ray.traceType = "py"
    rp = ray.Ray(-3.0, -9.5, 0, 1, freq, 1, n_field=nf)
    ray.Ray.traceType = "c"
    assert np.array_equal(rc.x, rp.x) and np.array_equal(rc.ky, rp.ky), "C and Python twins differ"
    # the raw grid-unit call, exactly as ray.py:59-83 prepares it
    c = 299792458.0 * 100
    w = 2 * np.pi * freq
    refr = nf.refractive_index(freq=freq, x=-3.0, y=-9.5)
    kx0 = float(np.asarray(0.0 * w / c * refr).reshape(-1)[0])
    ky0 = float(np.asarray(1.0 * w / c * refr).reshape(-1)[0])
    xg0 = float(nf.x2grid(-3.0)); yg0 = float(nf.y2grid(-9.5))
    ref = oracle.ref()
    L = 100000
    i, gx, gy, gkx, gky = run_ref_trace(ref, L, xg0, yg0, kx0, ky0, nf.st, nf.dx[0], nf.dy[0], nf.nx, nf.ny, nf.ax, nf.ay)
    assert i == len(rc.x)
    assert np.array_equal(gkx[:i], rc.kx) and np.array_equal(gky[:i], rc.ky), "launch-state restatement differs from Ray"
    # the class-level step multiplier (ray.py:12, :75): the same ray with Ray.kStep = 2 and 0.5
    ks = {}
    for tag, kstep in (("k2", 2), ("khalf", 0.5)):
        ray.Ray.kStep = kstep
        rk = ray.Ray(-3.0, -9.5, 0, 1, freq, 1, n_field=nf)
        ks.update({f"{tag}_kstep": np.float64(kstep), f"{tag}_n": np.int64(len(rk.x)), f"{tag}_x": rk.x, f"{tag}_y": rk.y,
                   f"{tag}_ph_end": np.float64(rk.ph[-1]), f"{tag}_st0": np.float64(rk.st[0])})
    ray.Ray.kStep = 1
    np.savez_compressed(
        os.path.join(HERE, "ft2_ray.npz"), **ks,
        nx=np.int64(nx), n0=4e13, edge=1e12, r_dia=7.85, half=10.0,
        field_freq=136.26e9, freq=freq, x0=-3.0, y0=-9.5,
        st=np.float64(nf.st), dx0=np.float64(nf.dx[0]), dy0=np.float64(nf.dy[0]),
        ax_sha=sha(nf.ax), ay_sha=sha(nf.ay), ax_sample=nf.ax[::20, ::20].copy(), ay_sample=nf.ay[::20, ::20].copy(),
        launch=np.array([xg0, yg0, kx0, ky0]), refr=np.float64(np.asarray(refr).reshape(-1)[0]),
        nsteps=np.int64(i), grid_traj=np.stack([gx[:i + 1], gy[:i + 1], gkx[:i + 1], gky[:i + 1]], 1),
        ray_x=rc.x, ray_y=rc.y, ray_kx=rc.kx, ray_ky=rc.ky, ray_st=rc.st, ray_lax=rc.lax, ray_ph=rc.ph, ray_k=rc.k)
    print("ft2_ray: n =", i, "end =", rc.x[-1], rc.y[-1])


def make_ft2_fan(den):
    ref = oracle.ref()
    out = {}
    rng = np.random.default_rng(20260101)
    nray = 96
    L = 4096
    for tag, nx, n0 in (("under", 201, 4e13), ("over", 129, 3.0e14)):
        n, x, y = ft2_density(nx, n0)
        nf = den.Nfield(n, x=x, y=y)
        # fan: start points on y = -9.5 cm, directions within +-0.35 rad of +y, |k| = w/c * N(edge)
        freq = 136.268e9
        c = 299792458.0 * 100
        k0 = 2 * np.pi * freq / c * float(nf.refractive_index(freq=freq))
        xs = rng.uniform(-6.0, 6.0, nray)
        ang = rng.uniform(-0.35, 0.35, nray)
        xg0 = np.asarray(nf.x2grid(xs), np.float64)
        yg0 = np.full(nray, float(nf.y2grid(-9.5)))
        kx0 = k0 * np.sin(ang)
        ky0 = k0 * np.cos(ang)
        nst = np.zeros(nray, np.int64)
        end = np.zeros((4, nray))
        dig = []
        ymax = np.zeros(nray)
        for r in range(nray):
            i, gx, gy, gkx, gky = run_ref_trace(ref, L, xg0[r], yg0[r], kx0[r], ky0[r], nf.st, nf.dx[0], nf.dy[0],
                                                nf.nx, nf.ny, nf.ax, nf.ay)
            nst[r] = i
            end[:, r] = gx[i], gy[i], gkx[i], gky[i]
            dig.append(sha(np.stack([gx[:i + 1], gy[:i + 1], gkx[:i + 1], gky[:i + 1]], 1)))
            ymax[r] = gy[:i + 1].max()
        out.update({
            f"{tag}_nx": np.int64(nx), f"{tag}_n0": n0, f"{tag}_st": np.float64(nf.st),
            f"{tag}_dx0": np.float64(nf.dx[0]), f"{tag}_dy0": np.float64(nf.dy[0]),
            f"{tag}_ax_sha": sha(nf.ax), f"{tag}_ay_sha": sha(nf.ay),
            f"{tag}_xg0": xg0, f"{tag}_yg0": yg0, f"{tag}_kx0": kx0, f"{tag}_ky0": ky0,
            f"{tag}_nsteps": nst, f"{tag}_end": end, f"{tag}_traj_sha": np.array(dig), f"{tag}_ymax": ymax,
        })
        print(f"ft2_fan[{tag}]: steps min/median/max = {nst.min()}/{int(np.median(nst))}/{nst.max()}",
              "reflected:", int((end[3] < 0).sum()))
    out["L"] = np.int64(L)
    out["edge"] = 1e12
    out["r_dia"] = 7.85
    out["half"] = 10.0
    np.savez_compressed(os.path.join(HERE, "ft2_fan.npz"), **out)


def make_rxyten_fan(den):
    """BASELINE config 3's density kind: shaped flux surfaces from the reference's own parameterisation
    (rxyten_from_params, den.py:458-485: Shafranov shift, triangularity, elongation), mapped to a
    Cartesian grid by RXYTEN.on_grid with 2-D meshgrids (den.py:169-230), NaNs outside the last surface
    filled with the edge density, then den.Nfield and a 64-ray fan through the compiled reference.
    The density grid itself is stored (97^2): the parameterisation is host prep outside the hot path."""
    ref = oracle.ref()
    rx, rlcs, _ = den.rxyten_from_params(shifx=0.3, delta=0.5, trian=0.1, elon=1.2, num=60)
    rx.den = 2.6e14 * np.maximum(0.0, 1 - (rx.r / abs(rlcs)) ** 2) + 1e12
    nx = 97
    x = np.linspace(-10, 10, nx)
    XX, YY = np.meshgrid(x, x)
    n = np.array(rx.on_grid(x=XX, y=YY)[0], dtype=np.float64)
    n[np.isnan(n)] = 1e12
    nf = den.Nfield(n, x=x, y=x)
    rng = np.random.default_rng(77)
    nray, L = 64, 4096
    freq = 136.268e9
    k0 = 2 * np.pi * freq / (299792458.0 * 100) * float(nf.refractive_index(freq=freq))
    xs = rng.uniform(-5.0, 5.0, nray)
    ang = rng.uniform(-0.3, 0.3, nray)
    xg0 = np.asarray(nf.x2grid(xs), np.float64)
    yg0 = np.full(nray, float(nf.y2grid(-9.5)))
    kx0, ky0 = k0 * np.sin(ang), k0 * np.cos(ang)
    nst = np.zeros(nray, np.int64)
    end = np.zeros((4, nray))
    dig = []
    for r in range(nray):
        i, gx, gy, gkx, gky = run_ref_trace(ref, L, xg0[r], yg0[r], kx0[r], ky0[r], nf.st, nf.dx[0], nf.dy[0],
                                            nf.nx, nf.ny, nf.ax, nf.ay)
        nst[r] = i
        end[:, r] = gx[i], gy[i], gkx[i], gky[i]
        dig.append(sha(np.stack([gx[:i + 1], gy[:i + 1], gkx[:i + 1], gky[:i + 1]], 1)))
    np.savez_compressed(os.path.join(HERE, "rxyten_fan.npz"), nx=np.int64(nx), density=n, x=x, rlcs=np.float64(rlcs),
                        st=np.float64(nf.st), ax_sha=sha(nf.ax), ay_sha=sha(nf.ay), xg0=xg0, yg0=yg0, kx0=kx0, ky0=ky0,
                        nsteps=nst, end=end, traj_sha=np.array(dig), L=np.int64(L))
    print("rxyten_fan: rlcs", float(rlcs), "steps min/median/max", nst.min(), int(np.median(nst)), nst.max(),
          "reflected", int((end[3] < 0).sum()), "density asymmetry", float(np.abs(n - n[:, ::-1]).max() / n.max()))


def make_ft2_cross(den, ray):
    """Reference Ray.addpoint (ray.py:131-224) on a 40-ray fan through the over-dense profile:
    what it returns for receiver lines y = 9, -2, -9 and x = 5 (first crossing)."""
    nx, n0 = 129, 3.0e14
    n, x, y = ft2_density(nx, n0)
    nf = den.Nfield(n, x=x, y=y)
    freq = 136.268e9
    rng = np.random.default_rng(77)
    nray = 40
    xs = rng.uniform(-6.0, 6.0, nray)
    ang = rng.uniform(-0.35, 0.35, nray)
    ray.Ray.MaxStep = 4096
    lines = [(9.0, "y"), (-2.0, "y"), (-9.0, "y"), (5.0, "x")]
    launch = np.zeros((nray, 4))
    nkept = np.zeros(nray, np.int64)
    res = np.full((len(lines), nray, 9), np.nan)      # count, index, x, y, kx, ky, lax, ph, raised
    from scipy.interpolate import interp1d
    look_t = np.full((len(lines), nray, 2), np.nan)   # target phases handed to the equal-phase lookup
    look_xy = np.full((len(lines), nray, 4), np.nan)  # interp1d(r.ph, r.x / r.y, extrapolate)(t) AFTER addpoint (beam.py:410-413)
    for r in range(nray):
        for li, (val, axis) in enumerate(lines):
            rr = ray.Ray(xs[r], -9.5, np.sin(ang[r]), np.cos(ang[r]), freq, 1, n_field=nf)
            launch[r] = float(nf.x2grid(xs[r])), float(nf.y2grid(-9.5)), rr.kx[0], rr.ky[0]
            nkept[r] = len(rr.x)
            try:
                ind = rr.addpoint(val, axis=axis)
            except ValueError:                        # empty .min(): no non-negative root
                res[li, r, 0] = 0; res[li, r, 8] = 1
                continue
            if not hasattr(ind, "__len__") or len(ind) == 0:
                res[li, r, 0] = 0; res[li, r, 8] = 0
                continue
            i = int(ind[0])
            res[li, r] = len(ind), i, rr.x[i], rr.y[i], rr.kx[i], rr.ky[i], rr.lax[i], rr.ph[i], 0
            if len(ind) == 1:
                ph = rr.ph
                t = np.array([(0.35 + 0.6 * rng.random()) * ph[i], (1.0 + 0.3 * rng.random()) * ph[i] + 1.0])
                if r % 7 == 0:
                    t[0] = -2.0                       # before the first node: first-segment extrapolation
                look_t[li, r] = t
                look_xy[li, r] = (interp1d(ph, rr.x, fill_value="extrapolate")(t[0]), interp1d(ph, rr.y, fill_value="extrapolate")(t[0]),
                                  interp1d(ph, rr.x, fill_value="extrapolate")(t[1]), interp1d(ph, rr.y, fill_value="extrapolate")(t[1]))
    ray.Ray.MaxStep = 100000
    np.savez_compressed(os.path.join(HERE, "ft2_cross.npz"), nx=np.int64(nx), n0=n0, freq=freq, x0=xs, y0=np.full(nray, -9.5),
                        launch=launch, nkept=nkept, st=np.float64(nf.st), dx0=np.float64(nf.dx[0]), dy0=np.float64(nf.dy[0]),
                        line_val=np.array([l[0] for l in lines]), line_axis=np.array([l[1] for l in lines]), res=res,
                        look_t=look_t, look_xy=look_xy)
    print("ft2_cross: counts per line", [(res[li, :, 0]).astype(int).tolist().count(1) for li in range(len(lines))],
          "none", [int((res[li, :, 0] == 0).sum()) for li in range(len(lines))],
          "multi", [int((res[li, :, 0] > 1).sum()) for li in range(len(lines))])


def make_beam(den):
    """Reference Beam (beam.py) on the FT-2 profile: launch conditions (Beam.rays, beam.py:278-344),
    field on the receiver line (Beam.field_r, beam.py:347-444) and the interferometer signal
    (field_result, beam.py:447-472).  All reference code, including its wavefront finder
    (beam.py:209-276) on top of the restated skimage.measure.find_contours of oracle/refenv.py; the
    front's nodes are stored so that tests can rebuild the identical interp1d."""
    import beam as refbeam
    nx, n0 = 201, 4e13
    n, x, y = ft2_density(nx, n0)
    nf = den.Nfield(n, x=x, y=y)
    x0 = 2.3
    y0 = -np.sqrt(100 - x0 ** 2)                      # antenna on the r = 10 cm circle (ft2.py:12-16)
    rnum = 101
    tx = refbeam.Beam(x0=x0, y0=y0, kx0=0, ky0=1)
    rx = refbeam.Beam(x0=x0, y0=-y0, kx0=0, ky0=-1)
    wf = tx.wavefront_r(dist=3, ampbnd=1e-3, n_field=nf)          # the reference's own finder
    lx, ly, lkx, lky, amps, phase, dist = tx.rays(rnum=rnum, dist=3, ampbnd=1e-3, n_field=nf)
    nkept = np.array([len(r.x) for r in tx._rays])
    field, xo, x_max, ph_max, signal, props = refbeam.field_result(tx, rx)
    # analytic pieces at a few points
    px = np.linspace(-1.5, 4.0, 7); py = np.linspace(-9.0, -4.0, 7)
    fs = tx.field(x=px, y=py, n=0.97, outfmt="dict")
    # tilted launches (the launch-angle axis of BASELINE config 4, Beam.k0 setter beam.py:96-100)
    tilt = {}
    for tag, th in (("tiltp", 0.12), ("tiltm", -0.2)):
        t2 = refbeam.Beam(x0=x0, y0=y0, kx0=0, ky0=1)
        t2.k0 = (np.sin(th), np.cos(th))
        r2 = refbeam.Beam(x0=x0, y0=-y0, kx0=0, ky0=-1)
        l2 = t2.rays(rnum=61, dist=3, ampbnd=1e-3, n_field=nf)
        nk2 = np.array([len(r.x) for r in t2._rays])          # before field_r inserts the receiver points
        f2, xo2, xm2, pm2, sg2, _ = refbeam.field_result(t2, r2)
        tilt.update({f"{tag}_angle": th, f"{tag}_launch": np.stack(l2[:6]), f"{tag}_nkept": nk2,
                     f"{tag}_field": f2, f"{tag}_xout": xo2, f"{tag}_x_max": xm2, f"{tag}_ph_max": np.float64(pm2),
                     f"{tag}_signal": np.complex128(sg2)})
    # ray counts that exercise the fine-grid normals of Beam.rays (beam.py:306-326, M = lcm(rnum, 1000)):
    # coprime (7 -> M = 7000), a divisor (250 -> stride 4), equal (1000 -> stride 1)
    for rn in (7, 250, 1000):
        t3 = refbeam.Beam(x0=x0, y0=y0, kx0=0, ky0=1)
        tilt[f"launch_r{rn}"] = np.stack(t3.rays(rnum=rn, dist=3, ampbnd=1e-3, n_field=nf)[:6])
    # class-level switches of the wavefront finder (beam.py:23, :27): no Gouy phase, a coarser phase map
    for tag, gouy, nfront in (("gouy0", 0, 1000), ("nfront500", 1, 500)):
        refbeam.Beam._gouy, refbeam.Beam._nfront = gouy, nfront
        w2 = refbeam.Beam(x0=x0, y0=y0, kx0=0, ky0=1).wavefront_r(dist=3, ampbnd=1e-3, n_field=nf)
        tilt.update({f"{tag}_wf_x": w2.x, f"{tag}_wf_y": w2.y})
    refbeam.Beam._gouy, refbeam.Beam._nfront = 1, 1000
    # other probing frequencies (the frequency axis of BASELINE config 4): field, step and beam all at f
    for tag, f in (("f100", 100e9), ("f160", 160e9)):
        nff = den.Nfield(n, x=x, y=y, freq=f)
        t2 = refbeam.Beam(x0=x0, y0=y0, kx0=0, ky0=1, freq=f)
        r2 = refbeam.Beam(x0=x0, y0=-y0, kx0=0, ky0=-1, freq=f)
        l2 = t2.rays(rnum=61, dist=3, ampbnd=1e-3, n_field=nff)
        nk2 = np.array([len(r.x) for r in t2._rays])
        f2, xo2, xm2, pm2, sg2, _ = refbeam.field_result(t2, r2)
        tilt.update({f"{tag}_freq": f, f"{tag}_st": np.float64(nff.st), f"{tag}_launch": np.stack(l2[:6]), f"{tag}_nkept": nk2,
                     f"{tag}_x_max": xm2, f"{tag}_ph_max": np.float64(pm2), f"{tag}_signal": np.complex128(sg2)})
    np.savez_compressed(os.path.join(HERE, "beam.npz"), nx=np.int64(nx), n0=n0, x0=x0, y0=y0, rnum=np.int64(rnum), **tilt,
                        wf_x=wf.x, wf_y=wf.y, launch=np.stack([lx, ly, lkx, lky, amps, phase]), nkept=nkept,
                        field=field, xout=xo, x_max=x_max, ph_max=np.float64(ph_max), signal=np.complex128(signal),
                        props=np.array(props if props is not None else [np.nan] * 3),
                        probe_x=px, probe_y=py, probe_field=fs["field"], probe_amp=fs["amp"], probe_phase=fs["phase"],
                        wave_len=tx.wave_len, rayleigh=tx.rayleigh_len(n=0.97))
    print("beam: rays kept", len(tx._rays), "of", rnum, "signal", signal, "ph_max", float(ph_max))


def make_ft2_phases():
    """The reference's top-level driver ft2.phases (ft2.py:25-71) end to end: density scan x seven
    antenna pairs -> interferometer phases.  Everything is reference code (Nfield, Beam incl. its
    wavefront finder, Ray, sidtrace.trace, addpoint, field_r, field_result); the only stand-in is the
    restated skimage.measure.find_contours of oracle/refenv.py."""
    import ft2 as refft2
    nx = 161
    n, x, y = ft2_density(nx, 4e13)
    a_set = np.array([0.0, 0.6, 1.0])
    ph_max, ph_sig = refft2.phases(n, x, y, a_set=a_set, nray=41)
    # strongly refracting plasmas (peak density 0.52 and 0.78 of critical): bent, partly folded beams.
    # (At a = 5.5 the reference itself fails with an IndexError in field_r; not a fixture.)
    dense = np.array([3.0, 4.5])
    dph_max, dph_sig = refft2.phases(n, x, y, a_set=dense, nray=41)
    np.savez_compressed(os.path.join(HERE, "ft2_phases.npz"), nx=np.int64(nx), n0=4e13, a_set=a_set, nray=np.int64(41),
                        ph_by_max=ph_max, ph_by_sig=ph_sig, dense_a_set=dense, dense_ph_by_max=dph_max,
                        dense_ph_by_sig=dph_sig)
    print("ft2_phases: ph_by_sig[a=1] =", np.round(ph_sig[-1], 4), " dense[a=4.5] =", np.round(dph_sig[-1], 4))


def ring_density(nx, half=10.0):
    """An over-dense ring around a hollow core: a closed cavity for rays launched inside it."""
    x = np.linspace(-half, half, nx)
    X, Y = np.meshgrid(x, x)
    r = np.sqrt(X ** 2 + Y ** 2)
    return 4.0e14 * np.exp(-((r - 6.0) / 2.5) ** 2) + 1e12, x, x


def make_long_rays(den):
    """Rays that need more than 2*nx steps -- the regime a fixed "a vacuum ray crosses the grid in ~nx
    steps" cap gets wrong: the reference's Beam.rays (one ray.Ray per start point, cap Ray.MaxStep =
    100000, ray.py:11) through a 201^2 FT-2 profile whose peak density sits at the cutoff (a narrow fan:
    the central rays crawl through N -> 0 for 500+ steps) and inside an over-dense ring (a cavity: every
    ray bounces between the walls for ~2000 steps, 10 x nx).  Stored per ray: launch conditions, number of
    kept points, last kept point, final eikonal phase and path length, SHA-256 of the four kept arrays."""
    import beam as refbeam
    out = {}
    nx = 201
    nc = den.critical_den(136.268e9)
    cases = (("near", ft2_density(nx, 1.001 * nc), (0.0, -10.0), 41, 0.9),
             ("ring", ring_density(nx), (0.3, -2.0), 21, 1e-3))
    for tag, (n, x, y), (xa, ya), rnum, ampbnd in cases:
        nf = den.Nfield(n, x=x, y=y)
        tx = refbeam.Beam(x0=xa, y0=ya, kx0=0, ky0=1)
        lx, ly, lkx, lky, amps, phase, dist = tx.rays(rnum=rnum, dist=3, ampbnd=ampbnd, n_field=nf)
        rays = tx._rays
        nk = np.array([len(r.x) for r in rays])
        out.update({f"{tag}_ant": np.array([xa, ya]), f"{tag}_rnum": np.int64(rnum), f"{tag}_ampbnd": np.float64(ampbnd),
                    f"{tag}_launch": np.stack([lx, ly, lkx, lky, amps]), f"{tag}_nkept": nk,
                    f"{tag}_last": np.array([[r.x[-1], r.y[-1], r.kx[-1], r.ky[-1]] for r in rays]),
                    f"{tag}_k0": np.array([[r.kx[0], r.ky[0]] for r in rays]),
                    f"{tag}_ph_end": np.array([r.ph[-1] for r in rays]), f"{tag}_lax_end": np.array([r.lax[-1] for r in rays]),
                    f"{tag}_sha": np.array([sha(np.stack([r.x, r.y, r.kx, r.ky], 1)) for r in rays]),
                    f"{tag}_st": np.float64(nf.st)})
        print(f"long_rays[{tag}]: kept points min/median/max = {nk.min()}/{int(np.median(nk))}/{nk.max()}, "
              f"beyond 2*nx = 402: {(nk > 402).sum()} of {len(nk)}, beyond 511: {(nk > 511).sum()}")
    out["near_n0"] = np.float64(1.001 * nc)
    np.savez_compressed(os.path.join(HERE, "long_rays.npz"), nx=np.int64(nx), **out)


if __name__ == "__main__":
    if not oracle.have_ref():
        oracle.build(ref=True)
    den, ray = refenv.modules()
    if "--only-long" in sys.argv:
        make_long_rays(den)
        sys.exit(0)
    make_kat(oracle.ref())
    make_ft2_ray(den, ray)
    make_ft2_fan(den)
    make_rxyten_fan(den)
    make_ft2_cross(den, ray)
    make_beam(den)
    make_ft2_phases()
    make_long_rays(den)
    for f in ("kat.npz", "ft2_ray.npz", "ft2_fan.npz", "rxyten_fan.npz", "ft2_cross.npz", "beam.npz", "ft2_phases.npz", "long_rays.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
