/* rk_oracle.c -- CPU checker of the adaptive-RK physics mode (include/sidtrace_b200_rk.h).
 *
 * TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use it; the product (rayrtrace_b200/) never does.
 *
 * PARITY UNPINNED.  The reference (AVSidorov/RAYRTRACE) has no adaptive integrator, no
 * interpolated field and no magnetised dispersion -- its hot path is sidtracemodule.c:36-48 only
 * (restated in sidtrace_oracle.c).  This file restates the algorithm of the EXTENSION
 *  on the CPU so that the CUDA kernel can be checked bit for bit (same operation
 * order, explicit fma, gcc -ffp-contract=off); the algorithm itself is validated by analytic
 * solutions and an independent SciPy integration in tests/test_rk_oracle.py.
 *
 *   dr/dtau = p,  dp/dtau = grad N^2 / 2,  N^2 = 1 - X  (O)  or  1 - X(1-X)/(1-X-Y^2)  (X mode)
 *   X = n/n_c from a node table (the n of den.py:65-66), bicubic Catmull-Rom or bilinear;
 *   Dormand-Prince 5(4), FSAL, max-norm error control, controller 0.9 err^(-1/5) in [0.2, 5]
 *   (err^(-1/5) by a division-free Newton iteration from an exponent-indexed seed).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const double *X, *alpha;
    int nx, ny;
    double xorg, yorg, rdx, rdy, Y0R0, R0;
    int xmode;  /* 0: O mode, 1: X mode with Y = Y0 R0/(R0+x), 2: X mode with Y^2 from the node table Y2 */
    int interp; /* 0 bilinear, 1 Catmull-Rom, 2 cubic B-spline on a prefiltered (ny+2) x (nx+2) table */
    const double *Y2;
} rk_medium;

typedef struct {
    double rtol, atol_r, atol_p, h0, hmin, hmax, tau_max;
    int max_steps;
    int stop_axis;
    double stop_val, stop_tol;
    /* histogram */
    double *hist;
    int nbins;
    double xaxis, yaxis, r_rlcs, bins_per_rho;
} rk_ctl;

typedef struct {
    double x, y, px, py, tau, phi, sig, n2min, ham;
    int nacc, nrej, status;
    int turn;
    double tx, ty, tpx, tpy;
} rk_out;

static const double A[6][6] = {
    {1.0 / 5.0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0},
};
static const double E[7] = {71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0,
                            -1.0 / 40.0};

static void cr_weights(double t, double *w, double *d)
{
    const double tt = t * t;
    w[0] = t * fma(t, fma(t, -0.5, 1.0), -0.5);
    w[1] = fma(tt, fma(t, 1.5, -2.5), 1.0);
    w[2] = t * fma(t, fma(t, -1.5, 2.0), 0.5);
    w[3] = tt * fma(t, 0.5, -0.5);
    d[0] = fma(t, fma(t, -1.5, 2.0), -0.5);
    d[1] = t * fma(t, 4.5, -5.0);
    d[2] = fma(t, fma(t, -4.5, 4.0), 0.5);
    d[3] = t * fma(t, 1.5, -1.0);
}

/* uniform cubic B-spline weights (C2 interpolant once the table is prefiltered) */
static void bs_weights(double t, double *w, double *d)
{
    const double tt = t * t, a = 1.0 - t;
    w[0] = ((a * a) * a) * (1.0 / 6.0);
    w[1] = fma(tt, fma(t, 0.5, -1.0), 2.0 / 3.0);
    w[2] = fma(t, fma(t, fma(t, -0.5, 0.5), 0.5), 1.0 / 6.0);
    w[3] = (tt * t) * (1.0 / 6.0);
    d[0] = -0.5 * (a * a);
    d[1] = t * fma(t, 1.5, -2.0);
    d[2] = fma(t, fma(t, -1.5, 1.0), 0.5);
    d[3] = 0.5 * tt;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void table_eval(const double *T, int nx, int ny, int interp, int iu, int iv, double t, double s, double *f,
                       double *fu, double *fv)
{
    if (interp == 0) {
        const double *r0 = T + (long long)iv * nx + iu;
        const double a = r0[0], b = r0[1], c = r0[nx], d = r0[nx + 1];
        const double ba = b - a, dc = d - c;
        const double lo = fma(t, ba, a), hi = fma(t, dc, c);
        *f = fma(s, hi - lo, lo);
        *fu = fma(s, dc - ba, ba);
        *fv = hi - lo;
        return;
    }
    double wx[4], dx[4], wy[4], dy[4], rv[4], rd[4];
    if (interp == 1) {
        cr_weights(t, wx, dx);
        cr_weights(s, wy, dy);
    } else {
        bs_weights(t, wx, dx);
        bs_weights(s, wy, dy);
    }
    for (int j = 0; j < 4; ++j) {
        double v0, v1, v2, v3;
        if (interp == 1) {
            const double *row = T + (long long)clampi(iv - 1 + j, 0, ny - 1) * nx;
            v0 = row[clampi(iu - 1, 0, nx - 1)]; v1 = row[iu]; v2 = row[iu + 1]; v3 = row[clampi(iu + 2, 0, nx - 1)];
        } else {   /* padded table: coefficient (-1, -1) is element (0, 0) */
            const double *row = T + (long long)(iv + j) * (nx + 2) + iu;
            v0 = row[0]; v1 = row[1]; v2 = row[2]; v3 = row[3];
        }
        rv[j] = fma(wx[3], v3, fma(wx[2], v2, fma(wx[1], v1, wx[0] * v0)));
        rd[j] = fma(dx[3], v3, fma(dx[2], v2, fma(dx[1], v1, dx[0] * v0)));
    }
    *f = fma(wy[3], rv[3], fma(wy[2], rv[2], fma(wy[1], rv[1], wy[0] * rv[0])));
    *fu = fma(wy[3], rd[3], fma(wy[2], rd[2], fma(wy[1], rd[1], wy[0] * rd[0])));
    *fv = fma(dy[3], rv[3], fma(dy[2], rv[2], fma(dy[1], rv[1], dy[0] * rv[0])));
}

/* force = grad(N^2)/2, absorption coefficient, N^2 at (x, y) */
void rk_oracle_eval(const rk_medium *m, double xs, double ys, double x, double y, double *fx, double *fy, double *al,
                    double *n2)
{
    const double u = (x - m->xorg) * m->rdx, v = (y - m->yorg) * m->rdy;
    const int iu = (int)fmin(fmax(u, 0.0), (double)(m->nx - 2));
    const int iv = (int)fmin(fmax(v, 0.0), (double)(m->ny - 2));
    const double t = u - (double)iu, s = v - (double)iv;
    double F, Fu, Fv;
    table_eval(m->X, m->nx, m->ny, m->interp, iu, iv, t, s, &F, &Fu, &Fv);
    const double Xv = xs * F;
    const double Xx = xs * (Fu * m->rdx), Xy = xs * (Fv * m->rdy);
    if (!m->xmode) {
        *n2 = 1.0 - Xv;
        *fx = -0.5 * Xx;
        *fy = -0.5 * Xy;
    } else if (m->xmode == 2) {
        double G, Gu, Gv;
        table_eval(m->Y2, m->nx, m->ny, m->interp, iu, iv, t, s, &G, &Gu, &Gv);
        const double ys2 = ys * ys;
        const double Y2 = ys2 * G;
        const double Y2x = ys2 * (Gu * m->rdx), Y2y = ys2 * (Gv * m->rdy);
        const double D = (1.0 - Xv) - Y2;
        const double rD = 1.0 / D;
        const double g = Xv * (1.0 - Xv);
        *n2 = 1.0 - g * rD;
        const double rD2 = rD * rD;
        const double dX = -(fma(1.0 - 2.0 * Xv, D, g)) * rD2;
        const double dY2 = -g * rD2;
        *fx = 0.5 * fma(dY2, Y2x, dX * Xx);
        *fy = 0.5 * fma(dY2, Y2y, dX * Xy);
    } else {
        const double rR = 1.0 / (m->R0 + x);
        const double Yv = ys * (m->Y0R0 * rR);
        const double Y2 = Yv * Yv;
        const double D = (1.0 - Xv) - Y2;
        const double rD = 1.0 / D;
        const double g = Xv * (1.0 - Xv);
        *n2 = 1.0 - g * rD;
        const double rD2 = rD * rD;
        const double dX = -(fma(1.0 - 2.0 * Xv, D, g)) * rD2;
        const double dY2 = -g * rD2;
        const double Y2x = -2.0 * (Y2 * rR);
        *fx = 0.5 * fma(dY2, Y2x, dX * Xx);
        *fy = 0.5 * (dX * Xy);
    }
    if (m->alpha) {
        double a, au, av;
        table_eval(m->alpha, m->nx, m->ny, m->interp, iu, iv, t, s, &a, &au, &av);
        *al = a;
    } else {
        *al = 0.0;
    }
}

/* Y^2 = (omega_ce/omega)^2 on the nodes of the medium's grid from a tabulated equilibrium: the arithmetic of
 * equilibrium_y2_kernel (rk_kernels.cuh), operation for operation.  psi[nr][nz] (R-major, equilibrium.py:34),
 * bilinear (equilibrium.py:46-49); fpol[nf] on a uniform psi grid from simagx to sibdry. */
void rk_oracle_equilibrium_y2(const double *psi_tab, int nr, int nz, double rleft, double rdim, double zmid, double zdim,
                              const double *fpol, int nf, double simagx, double sibdry, double r0_cm, double z0_cm,
                              double ypt, int nx, int ny, double xorg, double yorg, double dx, double dy, double *y2)
{
    const double zlow = zmid - zdim / 2;
    const double rdr = (double)(nr - 1) / rdim, rdz = (double)(nz - 1) / zdim;
    const double rdpsi = (double)(nf - 1) / (sibdry - simagx);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const double R = (r0_cm + (xorg + dx * (double)i)) * 0.01;
            const double Z = ((yorg + dy * (double)j) - z0_cm) * 0.01;
            const double u = (R - rleft) * rdr, v = (Z - zlow) * rdz;
            int iu = (int)floor(u), iv = (int)floor(v);
            iu = iu > nr - 2 ? nr - 2 : iu; iu = iu < 0 ? 0 : iu;
            iv = iv > nz - 2 ? nz - 2 : iv; iv = iv < 0 ? 0 : iv;
            const double t = u - (double)iu, s = v - (double)iv;
            const double *q = psi_tab + (long long)iu * nz + iv;
            const double p00 = q[0], p01 = q[1], p10 = q[nz], p11 = q[nz + 1];
            const double lo = p00 + t * (p10 - p00), hi = p01 + t * (p11 - p01);
            const double psi = lo + s * (hi - lo);
            /* grad psi: central differences at the four corner nodes, interpolated bilinearly (continuous field) */
            double gR[2][2], gZ[2][2];
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const int ia = iu + a, ib = iv + b;
                    const int a0 = ia - 1 > 0 ? ia - 1 : 0, a1 = ia + 1 < nr - 1 ? ia + 1 : nr - 1;
                    const int b0 = ib - 1 > 0 ? ib - 1 : 0, b1 = ib + 1 < nz - 1 ? ib + 1 : nz - 1;
                    gR[a][b] = (psi_tab[(long long)a1 * nz + ib] - psi_tab[(long long)a0 * nz + ib]) * (a1 - a0 == 2 ? 0.5 * rdr : rdr);
                    gZ[a][b] = (psi_tab[(long long)ia * nz + b1] - psi_tab[(long long)ia * nz + b0]) * (b1 - b0 == 2 ? 0.5 * rdz : rdz);
                }
            const double rlo = gR[0][0] + t * (gR[1][0] - gR[0][0]), rhi = gR[0][1] + t * (gR[1][1] - gR[0][1]);
            const double psiR = rlo + s * (rhi - rlo);
            const double zlo = gZ[0][0] + t * (gZ[1][0] - gZ[0][0]), zhi = gZ[0][1] + t * (gZ[1][1] - gZ[0][1]);
            const double psiZ = zlo + s * (zhi - zlo);
            double w = (psi - simagx) * rdpsi;
            w = w > 0.0 ? w : 0.0;
            const double wmax = (double)(nf - 1);
            w = w < wmax ? w : wmax;
            int iw = (int)floor(w);
            iw = iw > nf - 2 ? nf - 2 : iw;
            const double fw = w - (double)iw;
            const double F = fpol[iw] + fw * (fpol[iw + 1] - fpol[iw]);
            const double rR = 1.0 / R;
            const double BR = -(psiZ * rR), BZ = psiR * rR, Bp = F * rR;
            const double B2 = (BR * BR + BZ * BZ) + Bp * Bp;
            y2[(long long)j * nx + i] = (ypt * ypt) * B2;
        }
}

/* Cubic B-spline prefilter of one line (natural end conditions), the arithmetic of
 * bspline_prefilter_kernel: c[-1..n] from f[0..n-1]; out[0] holds c[-1]. */
static void prefilter_line(const double *f, long long fs, double *out, long long os, const double *R, int n)
{
    double *c = out + os;
    const double f0 = f[0], fl = f[(long long)(n - 1) * fs];
    c[0] = f0;
    c[(long long)(n - 1) * os] = fl;
    if (n > 2) {
        double dprev = 0.0;
        for (int i = 1; i <= n - 2; ++i) {
            double rhs = 6.0 * f[(long long)i * fs];
            if (i == 1) rhs -= f0;
            if (i == n - 2) rhs -= fl;
            const double d = (rhs - dprev) * R[i];
            c[(long long)i * os] = d;
            dprev = d;
        }
        double cnext = c[(long long)(n - 2) * os];
        for (int i = n - 3; i >= 1; --i) {
            const double ci = c[(long long)i * os] - R[i] * cnext;
            c[(long long)i * os] = ci;
            cnext = ci;
        }
    }
    c[-os] = 2.0 * c[0] - c[os];
    c[(long long)n * os] = 2.0 * c[(long long)(n - 1) * os] - c[(long long)(n - 2) * os];
}

/* nodes [ny][nx] -> coefficients [ny+2][nx+2] */
void rk_oracle_prefilter(const double *nodes, int nx, int ny, double *coef)
{
    const int nmax = nx > ny ? nx : ny;
    double *R = (double *)calloc((size_t)nmax + 1, sizeof(double));
    double *half = (double *)calloc((size_t)(nx + 2) * (size_t)ny, sizeof(double));
    R[1] = 0.25;
    for (int i = 2; i <= nmax; ++i) R[i] = 1.0 / (4.0 - R[i - 1]);
    for (int j = 0; j < ny; ++j) prefilter_line(nodes + (long long)j * nx, 1, half + (long long)j * (nx + 2), 1, R, nx);
    for (int i = 0; i < nx + 2; ++i) prefilter_line(half + i, nx + 2, coef + i, nx + 2, R, ny);
    free(R); free(half);
}

/* seeds 2^(-(k + 1/2)/5), k = -14 .. 10 (the bits of the kernel's cR5) */
static const double R5[25] = {
    0x1.9fdf8bcce533ep+2, 0x1.6a09e667f3bcdp+2, 0x1.3b2c47bff8328p+2, 0x1.125fbee250664p+2, 0x1.ddb680117ab12p+1,
    0x1.9fdf8bcce533dp+1, 0x1.6a09e667f3bcdp+1, 0x1.3b2c47bff8329p+1, 0x1.125fbee250664p+1, 0x1.ddb680117ab12p+0,
    0x1.9fdf8bcce533dp+0, 0x1.6a09e667f3bcdp+0, 0x1.3b2c47bff8329p+0, 0x1.125fbee250664p+0, 0x1.ddb680117ab12p-1,
    0x1.9fdf8bcce533ep-1, 0x1.6a09e667f3bcdp-1, 0x1.3b2c47bff8329p-1, 0x1.125fbee250664p-1, 0x1.ddb680117ab12p-2,
    0x1.9fdf8bcce533dp-2, 0x1.6a09e667f3bcdp-2, 0x1.3b2c47bff8329p-2, 0x1.125fbee250664p-2, 0x1.ddb680117ab12p-3};

static double inv_root5(double err)
{
    double e = err > 1e-4 ? err : 1e-4;
    e = e < 2e3 ? e : 2e3;
    uint64_t bits;
    memcpy(&bits, &e, sizeof bits);
    const int k = (int)((bits >> 52) & 0x7ff) - 1023 + 14;
    double r = R5[k];
    const double t = 0.2 * e;
    for (int it = 0; it < 4; ++it) {
        const double r2 = r * r;
        const double r5 = (r2 * r2) * r;
        r = r * fma(-t, r5, 1.2);
    }
    return r;
}

static double max2(double a, double b) { return a > b ? a : b; }
static double min2(double a, double b) { return a < b ? a : b; }

static int inside(const rk_medium *m, double x, double y)
{
    const double u = (x - m->xorg) * m->rdx, v = (y - m->yorg) * m->rdy;
    return u >= 0.0 && u < (double)(m->nx - 1) && v >= 0.0 && v < (double)(m->ny - 1);
}

/* One ray.  traj ([L][4]) and aux ([L][2]) may be NULL. */
void rk_oracle_ray(const rk_medium *m, const rk_ctl *c, double x, double y, double px, double py, double xs, double ys,
                   double amp, double *traj, double *aux, long long L, rk_out *o)
{
    double phi = 0.0, sig = 0.0, tau = 0.0, h = c->h0;
    const double amp2 = amp * amp;
    double P = amp2;
    int nacc = 0, nrej = 0, status, after_reject = 0;
    int turn = -1;
    double tx = NAN, ty = NAN, tpx = NAN, tpy = NAN;
    int land = 0, side = 0;
    double hland = 0, hl = 0, dl = 0, hu = 0, du = 0;
    double f1x, f1y, a1, n2, n2min;
    rk_oracle_eval(m, xs, ys, x, y, &f1x, &f1y, &a1, &n2);
    n2min = n2;
#define STORE()                                                                  \
    do {                                                                         \
        if (traj && nacc < L) {                                                  \
            double *tp = traj + (long long)nacc * 4;                             \
            tp[0] = x; tp[1] = y; tp[2] = px; tp[3] = py;                        \
            if (aux) { aux[(long long)nacc * 2] = tau; aux[(long long)nacc * 2 + 1] = sig; } \
        }                                                                        \
    } while (0)
    STORE();
    if (!inside(m, x, y)) { status = 4; goto done; }
    if (c->max_steps <= 0) { status = 2; goto done; }
    for (;;) {
        double hh = land ? hland : h;
        int last = 0;
        if (tau + hh >= c->tau_max) { hh = c->tau_max - tau; last = 1; }
        double kpx[7], kpy[7], kfx[7], kfy[7], ka[7];
        kpx[0] = px; kpy[0] = py; kfx[0] = f1x; kfy[0] = f1y; ka[0] = a1;
        double xn = x, yn = y, pxn = px, pyn = py, n2n = 0.0;
        for (int i = 0; i < 6; ++i) {
            double sx = A[i][0] * kpx[0], sy = A[i][0] * kpy[0], sfx = A[i][0] * kfx[0], sfy = A[i][0] * kfy[0];
            for (int j = 1; j <= i; ++j) {
                sx = fma(A[i][j], kpx[j], sx);
                sy = fma(A[i][j], kpy[j], sy);
                sfx = fma(A[i][j], kfx[j], sfx);
                sfy = fma(A[i][j], kfy[j], sfy);
            }
            xn = fma(hh, sx, x); yn = fma(hh, sy, y);
            pxn = fma(hh, sfx, px); pyn = fma(hh, sfy, py);
            kpx[i + 1] = pxn; kpy[i + 1] = pyn;
            rk_oracle_eval(m, xs, ys, xn, yn, &kfx[i + 1], &kfy[i + 1], &ka[i + 1], &n2n);
        }
        double ex = E[0] * kpx[0], ey = E[0] * kpy[0], efx = E[0] * kfx[0], efy = E[0] * kfy[0];
        for (int j = 1; j < 7; ++j) {
            ex = fma(E[j], kpx[j], ex);
            ey = fma(E[j], kpy[j], ey);
            efx = fma(E[j], kfx[j], efx);
            efy = fma(E[j], kfy[j], efy);
        }
        const double ah = fabs(hh);
        const double scx = fma(c->rtol, max2(fabs(x), fabs(xn)), c->atol_r);
        const double scy = fma(c->rtol, max2(fabs(y), fabs(yn)), c->atol_r);
        const double scpx = fma(c->rtol, max2(fabs(px), fabs(pxn)), c->atol_p);
        const double scpy = fma(c->rtol, max2(fabs(py), fabs(pyn)), c->atol_p);
        const double q0 = ah * fabs(ex), q1 = ah * fabs(ey), q2 = ah * fabs(efx), q3 = ah * fabs(efy);
        const double esum = (q0 + q1) + (q2 + q3);
        double en = q0, ed = scx;
        if (q1 * ed > en * scy) { en = q1; ed = scy; }
        if (q2 * ed > en * scpx) { en = q2; ed = scpx; }
        if (q3 * ed > en * scpy) { en = q3; ed = scpy; }
        const double err = (esum == esum) ? en / ed : esum;
        if (!(err <= 1.0)) {
            ++nrej;
            const double fac = (err == err) ? max2(0.2, 0.9 * inv_root5(err)) : 0.2;
            h = hh * fac;
            after_reject = 1;
            land = 0;
            if (!(h >= c->hmin)) { status = 3; goto done; }
            continue;
        }
        int on_line = 0;
        if (c->stop_axis) {
            const double d0 = (c->stop_axis == 1 ? y : x) - c->stop_val;
            const double d1 = (c->stop_axis == 1 ? yn : xn) - c->stop_val;
            if (!land) {
                if (d0 != 0.0 && ((d0 < 0.0) != (d1 < 0.0) || d1 == 0.0)) {
                    if (fabs(d1) <= c->stop_tol) {
                        on_line = 1;
                    } else {
                        land = 1; side = 0;
                        hl = 0.0; dl = d0; hu = hh; du = d1;
                        hland = hh * (d0 / (d0 - d1));
                        ++nrej;
                        continue;
                    }
                }
            } else if (fabs(d1) <= c->stop_tol || land >= 12) {
                on_line = 1;
            } else {
                if ((d1 < 0.0) == (dl < 0.0)) {
                    hl = hh; dl = d1;
                    if (side == -1) du *= 0.5;
                    side = -1;
                } else {
                    hu = hh; du = d1;
                    if (side == 1) dl *= 0.5;
                    side = 1;
                }
                hland = (hl * du - hu * dl) / (du - dl);
                ++land; ++nrej;
                continue;
            }
        }
        double sphi = A[5][0] * fma(kpx[0], kpx[0], kpy[0] * kpy[0]);
        double ssig = A[5][0] * ka[0];
        for (int j = 1; j < 6; ++j) {
            sphi = fma(A[5][j], fma(kpx[j], kpx[j], kpy[j] * kpy[j]), sphi);
            ssig = fma(A[5][j], ka[j], ssig);
        }
        const double xm = 0.5 * (x + xn), ym = 0.5 * (y + yn);
        if (turn < 0 && ((py > 0.0 && pyn <= 0.0) || (py < 0.0 && pyn >= 0.0))) {
            turn = nacc + 1;
            tx = xn; ty = yn; tpx = pxn; tpy = pyn;
        }
        x = xn; y = yn; px = pxn; py = pyn;
        phi = fma(hh, sphi, phi);
        sig = fma(hh, ssig, sig);
        tau = last ? c->tau_max : tau + hh;
        f1x = kfx[6]; f1y = kfy[6]; a1 = ka[6];
        n2min = min2(n2n, n2min);
        ++nacc;
        STORE();
        if (c->hist && m->alpha) {
            const double Pn = amp2 * exp(-sig);
            const double dP = P - Pn;
            P = Pn;
            const double rx = xm - c->xaxis, ry = ym - c->yaxis;
            const double rho = sqrt(fma(rx, rx, ry * ry)) * c->r_rlcs;
            const double fb = rho * c->bins_per_rho;
            if (fb < (double)c->nbins) c->hist[(int)fb] += dP;
        }
        /* no growth right after a rejection: keeps the controller from bouncing at the kinks of
         * the interpolant (cell edges) */
        const double fmx = after_reject ? 1.0 : 5.0;
        const double fac = min2(0.9 * inv_root5(err), fmx);
        after_reject = 0;
        h = min2(hh * fac, c->hmax);
        if (on_line) { status = 5; break; }
        if (!inside(m, x, y)) { status = 0; break; }
        if (last) { status = 1; break; }
        if (nacc >= c->max_steps) { status = 2; break; }
    }
done:
#undef STORE
    o->x = x; o->y = y; o->px = px; o->py = py;
    o->tau = tau; o->phi = phi; o->sig = sig; o->n2min = n2min;
    {
        double fx, fy, al, n2e;
        rk_medium mm = *m;
        mm.alpha = NULL;
        rk_oracle_eval(&mm, xs, ys, x, y, &fx, &fy, &al, &n2e);
        o->ham = fma(px, px, py * py) - n2e;
    }
    o->nacc = nacc; o->nrej = nrej; o->status = status;
    o->turn = turn; o->tx = tx; o->ty = ty; o->tpx = tpx; o->tpy = tpy;
}

/* ---- batch driver (pthreads; per-thread histograms summed in thread order) ---- */
typedef struct {
    const rk_medium *m;
    rk_ctl c;
    long long lo, hi, n, L;
    const double *x0, *y0, *px0, *py0, *xscale, *yscale, *amp;
    double *end_state, *tau, *phi, *sig, *n2min, *ham, *traj, *aux;
    int32_t *nacc, *nrej, *status, *turn;
    double *turn_state;
} rk_job;

static void *rk_worker(void *arg)
{
    rk_job *j = (rk_job *)arg;
    for (long long r = j->lo; r < j->hi; ++r) {
        rk_out o;
        rk_oracle_ray(j->m, &j->c, j->x0[r], j->y0[r], j->px0[r], j->py0[r], j->xscale ? j->xscale[r] : 1.0,
                      j->yscale ? j->yscale[r] : 1.0, j->amp ? j->amp[r] : 1.0, j->traj ? j->traj + r * j->L * 4 : NULL,
                      j->aux ? j->aux + r * j->L * 2 : NULL, j->L, &o);
        if (j->end_state) {
            j->end_state[r] = o.x; j->end_state[j->n + r] = o.y;
            j->end_state[2 * j->n + r] = o.px; j->end_state[3 * j->n + r] = o.py;
        }
        if (j->tau) j->tau[r] = o.tau;
        if (j->phi) j->phi[r] = o.phi;
        if (j->sig) j->sig[r] = o.sig;
        if (j->n2min) j->n2min[r] = o.n2min;
        if (j->ham) j->ham[r] = o.ham;
        if (j->nacc) j->nacc[r] = o.nacc;
        if (j->nrej) j->nrej[r] = o.nrej;
        if (j->status) j->status[r] = o.status;
        if (j->turn) j->turn[r] = o.turn;
        if (j->turn_state) {
            j->turn_state[r] = o.tx; j->turn_state[j->n + r] = o.ty;
            j->turn_state[2 * j->n + r] = o.tpx; j->turn_state[3 * j->n + r] = o.tpy;
        }
    }
    return NULL;
}

int rk_oracle_batch(const rk_medium *m, const rk_ctl *c, long long n, const double *x0, const double *y0,
                    const double *px0, const double *py0, const double *xscale, const double *yscale, const double *amp,
                    double *end_state, double *tau, double *phi, double *sig, double *n2min, double *ham, int32_t *nacc,
                    int32_t *nrej, int32_t *status, double *traj, double *aux, long long L, int nthreads, int32_t *turn,
                    double *turn_state)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (n < nthreads) nthreads = n > 0 ? (int)n : 1;
    rk_job *jobs = (rk_job *)calloc((size_t)nthreads, sizeof(rk_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    double *hists = NULL;
    if (c->hist) hists = (double *)calloc((size_t)nthreads * (size_t)c->nbins, sizeof(double));
    for (int t = 0; t < nthreads; ++t) {
        rk_job *j = &jobs[t];
        j->m = m; j->c = *c;
        if (c->hist) j->c.hist = hists + (size_t)t * (size_t)c->nbins;
        j->lo = n * t / nthreads; j->hi = n * (t + 1) / nthreads; j->n = n; j->L = L;
        j->x0 = x0; j->y0 = y0; j->px0 = px0; j->py0 = py0; j->xscale = xscale; j->yscale = yscale; j->amp = amp;
        j->end_state = end_state; j->tau = tau; j->phi = phi; j->sig = sig; j->n2min = n2min; j->ham = ham;
        j->nacc = nacc; j->nrej = nrej; j->status = status; j->traj = traj; j->aux = aux;
        j->turn = turn; j->turn_state = turn_state;
        if (nthreads == 1) rk_worker(j);
        else pthread_create(&th[t], NULL, rk_worker, j);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    if (c->hist) {
        for (int t = 0; t < nthreads; ++t)
            for (int b = 0; b < c->nbins; ++b) c->hist[b] += hists[(size_t)t * (size_t)c->nbins + b];
        free(hists);
    }
    free(jobs); free(th);
    return 0;
}
