// rk_kernels.cuh -- beyond-reference physics mode: adaptive Dormand-Prince 5(4)
// integration of the Hamiltonian ray equations
//
//     dr/dtau = p,   dp/dtau = (1/2) grad N^2(r),   p = k / k0,  H = |p|^2 - N^2(r) = 0
//
// through a density table X = n/n_c interpolated by a cubic B-spline (C2, prefiltered on the device),
// Catmull-Rom patches (C1, local) or bilinearly,
// with the cold-plasma O-mode (N^2 = 1 - X) or X-mode (N^2 = 1 - X(1-X)/(1-X-Y^2), toroidal
// Y = Y0 R0/(R0+x)) dispersion and its analytic derivatives fused into the field evaluation,
// an optional absorption table integrated as optical depth, and absorbed power deposited into
// flux-coordinate bins.  The reference has NO counterpart of this (its only integrator is the
// fixed-step loop of sidtracemodule.c:36-48, which trace_kernels.cuh reproduces bit for bit):
// this file is an extension and is never reported as parity.  Its checker is oracle/rk_oracle.c
// (same operation order, explicit fma -> bit-identical) plus analytic solutions in the tests.
//
// One thread per ray; rays take different numbers of (accepted + rejected) steps, so warps are
// persistent and refill idle lanes from a global queue after a vote, like trace_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stbrk {

constexpr int kBlock = 128;
constexpr unsigned FULL = 0xffffffffu;

enum Status : int { LEFT_DOMAIN = 0, TAU_MAX = 1, MAX_STEPS = 2, STEP_UNDERFLOW = 3, START_OUTSIDE = 4, STOP_LINE = 5 };

struct Medium {
    const double *X;       // n/n_c at the reference frequency: [ny][nx] nodes, or [ny+2][nx+2] B-spline coefficients
    const double *alpha;   // power absorption coefficient per unit tau (same layout), or nullptr
    const double2 *XA;     // {X, alpha} interleaved per node (media with absorption), or nullptr
    const double *Y2;      // MODE 2: (omega_ce/omega)^2 node table (same layout as X), or nullptr
    const double2 *XY;     // MODE 2 without absorption: {X, Y^2} interleaved per node (one 16-byte load serves both), or nullptr
    int nx, ny;
    double xorg, yorg;     // physical position of node (0, 0)
    double rdx, rdy;       // 1/dx, 1/dy
    double Y0R0, R0;       // X-mode: Y(x) = Y0R0 / (R0 + x)
};

struct Params {
    Medium m;
    long long nrays;
    const double *x0, *y0, *px0, *py0;   // launch position (cm) and refractive-index vector
    const double *xscale, *yscale;       // per-ray multipliers of X and Y (sweeps) or nullptr
    const double *amp;                   // launch amplitude or nullptr (-> 1)
    double rtol, atol_r, atol_p;
    double h0, hmin, hmax, tau_max;
    int max_steps;
    int refill_min;                      // idle lanes that trigger a refill: 32 = warp tiles ... 1 = lane refill
    int stop_axis;                       // 0: none, 1: stop on the line y = stop_val, 2: x = stop_val
    double stop_val, stop_tol;
    // outputs (any may be nullptr)
    double *end_state;                   // [4][nrays]
    double *tau_end, *phi_end, *sig_end, *n2_min, *ham_end;
    int *nacc, *nrej, *status;
    int *turn_index;                     // first accepted step at which p_y changed sign, or -1
    double *turn_state;                  // [4][nrays] the state there (NaN when none)
    double *traj;                        // [ray][L][4] accepted states (state 0 = launch)
    double *traj_aux;                    // [ray][L][2] (tau, sigma)
    long long L;
    double *hist;
    int nbins;
    double xaxis, yaxis, r_rlcs, bins_per_rho;
    unsigned long long *counter;
};

// Dormand-Prince 5(4) tableau.  cA[i][j] = a_{i+2, j+1}; row 5 is also the 5th-order weights
// (FSAL); cE = b - b* (error estimate).
__constant__ double cA[6][6] = {
    {1.0 / 5.0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0},
};
__constant__ double cE[7] = {71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0,
                             -1.0 / 40.0};

// Catmull-Rom weights and their derivatives at fractional position t (Horner, explicit fma)
__device__ __forceinline__ void cr_weights(double t, double (&w)[4], double (&d)[4])
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

// uniform cubic B-spline weights: a C2 interpolant once the table is prefiltered
__device__ __forceinline__ void bs_weights(double t, double (&w)[4], double (&d)[4])
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

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Table value and grid-unit gradient at grid position (u, v).  INTERP 0: bilinear, 1: Catmull-Rom on
// the node table, 2: cubic B-spline on the prefiltered (ny+2) x (nx+2) coefficient table (element
// (0, 0) is coefficient (-1, -1), so the 4 x 4 stencil never leaves it).
// The cell index is clamped to the table, the fractional part is not: outside the table the
// edge cell's polynomial is continued (stage points of the last step may fall there).
template <int INTERP, bool GRAD>
__device__ __forceinline__ void table_eval(const double *__restrict__ T, int nx, int ny, int iu, int iv, double t,
                                           double s, double &f, double &fu, double &fv)
{
    if (INTERP == 0) {
        const double *r0 = T + (long long)iv * nx + iu;
        const double a = __ldg(r0), b = __ldg(r0 + 1), c = __ldg(r0 + nx), d = __ldg(r0 + nx + 1);
        const double ba = b - a, dc = d - c;
        const double lo = fma(t, ba, a), hi = fma(t, dc, c);
        f = fma(s, hi - lo, lo);
        if (GRAD) {
            fu = fma(s, dc - ba, ba);
            fv = hi - lo;
        }
    } else {
        double wx[4], dx[4], wy[4], dy[4];
        if (INTERP == 1) {
            cr_weights(t, wx, dx);
            cr_weights(s, wy, dy);
        } else {
            bs_weights(t, wx, dx);
            bs_weights(s, wy, dy);
        }
        double rv[4], rd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double v0, v1, v2, v3;
            if (INTERP == 1) {
                const double *row = T + (long long)clampi(iv - 1 + j, 0, ny - 1) * nx;
                v0 = __ldg(row + clampi(iu - 1, 0, nx - 1)); v1 = __ldg(row + iu); v2 = __ldg(row + iu + 1);
                v3 = __ldg(row + clampi(iu + 2, 0, nx - 1));
            } else {
                const double *row = T + (long long)(iv + j) * (nx + 2) + iu;
                v0 = __ldg(row); v1 = __ldg(row + 1); v2 = __ldg(row + 2); v3 = __ldg(row + 3);
            }
            rv[j] = fma(wx[3], v3, fma(wx[2], v2, fma(wx[1], v1, wx[0] * v0)));
            if (GRAD) rd[j] = fma(dx[3], v3, fma(dx[2], v2, fma(dx[1], v1, dx[0] * v0)));
        }
        f = fma(wy[3], rv[3], fma(wy[2], rv[2], fma(wy[1], rv[1], wy[0] * rv[0])));
        if (GRAD) {
            fu = fma(wy[3], rd[3], fma(wy[2], rd[2], fma(wy[1], rd[1], wy[0] * rd[0])));
            fv = fma(dy[3], rv[3], fma(dy[2], rv[2], fma(dy[1], rv[1], dy[0] * rv[0])));
        }
    }
}

// The same for a medium with absorption: X and alpha interleaved per node ({X, alpha} as one double2), so
// that one 16-byte load serves both tables (separate tables cost 32 scattered loads per evaluation and
// twice the L1 footprint: measured 9.1 ms per 10^6-ray beam against 4.2 ms without absorption).
// Each table sees exactly the arithmetic of table_eval, so results do not depend on the layout.
template <int INTERP>
__device__ __forceinline__ void table_eval2(const double2 *__restrict__ T, int nx, int ny, int iu, int iv, double t,
                                            double s, double &f, double &fu, double &fv, double &al)
{
    if (INTERP == 0) {
        const double2 *r0 = T + (long long)iv * nx + iu;
        const double2 a = __ldg(r0), b = __ldg(r0 + 1), c = __ldg(r0 + nx), d = __ldg(r0 + nx + 1);
        const double ba = b.x - a.x, dc = d.x - c.x;
        const double lo = fma(t, ba, a.x), hi = fma(t, dc, c.x);
        f = fma(s, hi - lo, lo);
        fu = fma(s, dc - ba, ba);
        fv = hi - lo;
        const double lo2 = fma(t, b.y - a.y, a.y), hi2 = fma(t, d.y - c.y, c.y);
        al = fma(s, hi2 - lo2, lo2);
    } else {
        double wx[4], dx[4], wy[4], dy[4];
        if (INTERP == 1) {
            cr_weights(t, wx, dx);
            cr_weights(s, wy, dy);
        } else {
            bs_weights(t, wx, dx);
            bs_weights(s, wy, dy);
        }
        double rv[4], rd[4], ra[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 v0, v1, v2, v3;
            if (INTERP == 1) {
                const double2 *row = T + (long long)clampi(iv - 1 + j, 0, ny - 1) * nx;
                v0 = __ldg(row + clampi(iu - 1, 0, nx - 1)); v1 = __ldg(row + iu); v2 = __ldg(row + iu + 1);
                v3 = __ldg(row + clampi(iu + 2, 0, nx - 1));
            } else {
                const double2 *row = T + (long long)(iv + j) * (nx + 2) + iu;
                v0 = __ldg(row); v1 = __ldg(row + 1); v2 = __ldg(row + 2); v3 = __ldg(row + 3);
            }
            rv[j] = fma(wx[3], v3.x, fma(wx[2], v2.x, fma(wx[1], v1.x, wx[0] * v0.x)));
            rd[j] = fma(dx[3], v3.x, fma(dx[2], v2.x, fma(dx[1], v1.x, dx[0] * v0.x)));
            ra[j] = fma(wx[3], v3.y, fma(wx[2], v2.y, fma(wx[1], v1.y, wx[0] * v0.y)));
        }
        f = fma(wy[3], rv[3], fma(wy[2], rv[2], fma(wy[1], rv[1], wy[0] * rv[0])));
        fu = fma(wy[3], rd[3], fma(wy[2], rd[2], fma(wy[1], rd[1], wy[0] * rd[0])));
        fv = fma(dy[3], rv[3], fma(dy[2], rv[2], fma(dy[1], rv[1], dy[0] * rv[0])));
        al = fma(wy[3], ra[3], fma(wy[2], ra[2], fma(wy[1], ra[1], wy[0] * ra[0])));
    }
}

// Value and gradient of BOTH components of an interleaved table ({X, Y^2}, MODE 2): each component sees exactly the
// arithmetic of table_eval, so results do not depend on the layout.
template <int INTERP>
__device__ __forceinline__ void table_eval2g(const double2 *__restrict__ T, int nx, int ny, int iu, int iv, double t, double s,
                                             double &f, double &fu, double &fv, double &g, double &gu, double &gv)
{
    if (INTERP == 0) {
        const double2 *r0 = T + (long long)iv * nx + iu;
        const double2 a = __ldg(r0), b = __ldg(r0 + 1), c = __ldg(r0 + nx), d = __ldg(r0 + nx + 1);
        {
            const double ba = b.x - a.x, dc = d.x - c.x;
            const double lo = fma(t, ba, a.x), hi = fma(t, dc, c.x);
            f = fma(s, hi - lo, lo); fu = fma(s, dc - ba, ba); fv = hi - lo;
        }
        {
            const double ba = b.y - a.y, dc = d.y - c.y;
            const double lo = fma(t, ba, a.y), hi = fma(t, dc, c.y);
            g = fma(s, hi - lo, lo); gu = fma(s, dc - ba, ba); gv = hi - lo;
        }
    } else {
        double wx[4], dx[4], wy[4], dy[4];
        if (INTERP == 1) {
            cr_weights(t, wx, dx);
            cr_weights(s, wy, dy);
        } else {
            bs_weights(t, wx, dx);
            bs_weights(s, wy, dy);
        }
        double rv[4], rd[4], qv[4], qd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 v0, v1, v2, v3;
            if (INTERP == 1) {
                const double2 *row = T + (long long)clampi(iv - 1 + j, 0, ny - 1) * nx;
                v0 = __ldg(row + clampi(iu - 1, 0, nx - 1)); v1 = __ldg(row + iu); v2 = __ldg(row + iu + 1);
                v3 = __ldg(row + clampi(iu + 2, 0, nx - 1));
            } else {
                const double2 *row = T + (long long)(iv + j) * (nx + 2) + iu;
                v0 = __ldg(row); v1 = __ldg(row + 1); v2 = __ldg(row + 2); v3 = __ldg(row + 3);
            }
            rv[j] = fma(wx[3], v3.x, fma(wx[2], v2.x, fma(wx[1], v1.x, wx[0] * v0.x)));
            rd[j] = fma(dx[3], v3.x, fma(dx[2], v2.x, fma(dx[1], v1.x, dx[0] * v0.x)));
            qv[j] = fma(wx[3], v3.y, fma(wx[2], v2.y, fma(wx[1], v1.y, wx[0] * v0.y)));
            qd[j] = fma(dx[3], v3.y, fma(dx[2], v2.y, fma(dx[1], v1.y, dx[0] * v0.y)));
        }
        f = fma(wy[3], rv[3], fma(wy[2], rv[2], fma(wy[1], rv[1], wy[0] * rv[0])));
        fu = fma(wy[3], rd[3], fma(wy[2], rd[2], fma(wy[1], rd[1], wy[0] * rd[0])));
        fv = fma(dy[3], rv[3], fma(dy[2], rv[2], fma(dy[1], rv[1], dy[0] * rv[0])));
        g = fma(wy[3], qv[3], fma(wy[2], qv[2], fma(wy[1], qv[1], wy[0] * qv[0])));
        gu = fma(wy[3], qd[3], fma(wy[2], qd[2], fma(wy[1], qd[1], wy[0] * qd[0])));
        gv = fma(dy[3], qv[3], fma(dy[2], qv[2], fma(dy[1], qv[1], dy[0] * qv[0])));
    }
}

// Right-hand side at (x, y): force = grad(N^2)/2, absorption coefficient, N^2.
#ifndef STB_RK_EVAL_INLINE
#define STB_RK_EVAL_INLINE __forceinline__
#endif
#ifndef STB_RK_MINB
#define STB_RK_MINB 3   // <= 170 registers, 3 CTAs per SM: measured best (profiles/r01/rk_occupancy_sweep.txt)
#endif
// MODE: 0 = O mode, 1 = X mode with the analytic toroidal field Y = Y0 R0/(R0+x), 2 = X mode with Y^2 from a node
// table (|B| of a tabulated equilibrium, equilibrium_y2_kernel)
template <int INTERP, int MODE, bool ABSORB>
__device__ STB_RK_EVAL_INLINE void rk_eval(const Medium &m, double xs, double ys, double x, double y, double &fx,
                                        double &fy, double &al, double &n2)
{
    const double u = (x - m.xorg) * m.rdx, v = (y - m.yorg) * m.rdy;
    // cell index = clamp(floor(u), 0, nx-2): the saturating conversion (NaN -> 0) and integer min/max
    // give what (int)fmin(fmax(u, 0), nx-2) gives, without the NaN-aware double min/max sequences
    const int iu = max(0, min(__double2int_rd(u), m.nx - 2));
    const int iv = max(0, min(__double2int_rd(v), m.ny - 2));
    const double t = u - (double)iu, s = v - (double)iv;
    double F, Fu, Fv, G = 0, Gu = 0, Gv = 0;
    if (ABSORB) table_eval2<INTERP>(m.XA, m.nx, m.ny, iu, iv, t, s, F, Fu, Fv, al);
    else if (MODE == 2) table_eval2g<INTERP>(m.XY, m.nx, m.ny, iu, iv, t, s, F, Fu, Fv, G, Gu, Gv);
    else table_eval<INTERP, true>(m.X, m.nx, m.ny, iu, iv, t, s, F, Fu, Fv);
    const double Xv = xs * F;
    const double Xx = xs * (Fu * m.rdx), Xy = xs * (Fv * m.rdy);   // physical gradient of X
    if (MODE == 0) {
        n2 = 1.0 - Xv;
        fx = -0.5 * Xx;
        fy = -0.5 * Xy;
    } else if (MODE == 2) {
        if (ABSORB) table_eval<INTERP, true>(m.Y2, m.nx, m.ny, iu, iv, t, s, G, Gu, Gv);   // {X, alpha} came interleaved
        const double ys2 = ys * ys;
        const double Y2 = ys2 * G;
        const double Y2x = ys2 * (Gu * m.rdx), Y2y = ys2 * (Gv * m.rdy);   // physical gradient of Y^2
        const double D = (1.0 - Xv) - Y2;
        const double rD = 1.0 / D;
        const double g = Xv * (1.0 - Xv);
        n2 = 1.0 - g * rD;
        const double rD2 = rD * rD;
        const double dX = -(fma(1.0 - 2.0 * Xv, D, g)) * rD2;    // dN^2/dX
        const double dY2 = -g * rD2;                             // dN^2/d(Y^2)
        fx = 0.5 * fma(dY2, Y2x, dX * Xx);
        fy = 0.5 * fma(dY2, Y2y, dX * Xy);
    } else {
        const double rR = 1.0 / (m.R0 + x);
        const double Yv = ys * (m.Y0R0 * rR);
        const double Y2 = Yv * Yv;
        const double D = (1.0 - Xv) - Y2;
        const double rD = 1.0 / D;
        const double g = Xv * (1.0 - Xv);
        n2 = 1.0 - g * rD;
        const double rD2 = rD * rD;
        const double dX = -(fma(1.0 - 2.0 * Xv, D, g)) * rD2;    // dN^2/dX
        const double dY2 = -g * rD2;                             // dN^2/d(Y^2)
        const double Y2x = -2.0 * (Y2 * rR);                     // d(Y^2)/dx
        fx = 0.5 * fma(dY2, Y2x, dX * Xx);
        fy = 0.5 * (dX * Xy);
    }
    if (!ABSORB) al = 0.0;
}

// Seeds 2^(-(k + 1/2)/5), k = -14 .. 10, of the inverse fifth root (hex literals: the CPU checker holds
// the same bits).
__constant__ double cR5[25] = {
    0x1.9fdf8bcce533ep+2, 0x1.6a09e667f3bcdp+2, 0x1.3b2c47bff8328p+2, 0x1.125fbee250664p+2, 0x1.ddb680117ab12p+1,
    0x1.9fdf8bcce533dp+1, 0x1.6a09e667f3bcdp+1, 0x1.3b2c47bff8329p+1, 0x1.125fbee250664p+1, 0x1.ddb680117ab12p+0,
    0x1.9fdf8bcce533dp+0, 0x1.6a09e667f3bcdp+0, 0x1.3b2c47bff8329p+0, 0x1.125fbee250664p+0, 0x1.ddb680117ab12p-1,
    0x1.9fdf8bcce533ep-1, 0x1.6a09e667f3bcdp-1, 0x1.3b2c47bff8329p-1, 0x1.125fbee250664p-1, 0x1.ddb680117ab12p-2,
    0x1.9fdf8bcce533dp-2, 0x1.6a09e667f3bcdp-2, 0x1.3b2c47bff8329p-2, 0x1.125fbee250664p-2, 0x1.ddb680117ab12p-3};

// err^(-1/5) for the step-size controller, from IEEE operations only so that the CPU checker reproduces
// the step-size sequence bit for bit: seed from the binary exponent, then four division-free Newton
// steps r <- r (1.2 - 0.2 e r^5) (relative error ~1e-11; the controller needs far less).  The argument
// is clamped to the range the controller's [0.2, 5] limits can distinguish.
__device__ __forceinline__ double inv_root5(double err)
{
    double e = err > 1e-4 ? err : 1e-4;
    e = e < 2e3 ? e : 2e3;
    const int k = ((__double2hiint(e) >> 20) & 0x7ff) - 1023 + 14;
    double r = cR5[k];
    const double t = 0.2 * e;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const double r2 = r * r;
        const double r5 = (r2 * r2) * r;
        r = r * fma(-t, r5, 1.2);
    }
    return r;
}

// compare-select max / min (no NaN special-casing: NaNs are caught separately)
__device__ __forceinline__ double max2(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double min2(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ void rk_hist_add(double *sh, int b, double w)
{
    const unsigned peers = __match_any_sync(__activemask(), b);
    const int lane = threadIdx.x & 31;
    if (peers == FULL) {
        // the whole warp deposits into one bin (neighbouring rays of a beam): butterfly, one atomic
        double s = w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0 && b >= 0) atomicAdd(&sh[b], s);
        return;
    }
    // several bins (or a partly active warp): one atomic per distinct bin, lanes summed in lane order
    const int leader = __ffs(peers) - 1;
    double sum = 0.0;
    for (unsigned mm = peers; mm; mm &= mm - 1) sum += __shfl_sync(peers, w, __ffs(mm) - 1);
    if (lane == leader && b >= 0) atomicAdd(&sh[b], sum);
}

template <int INTERP, int MODE, bool ABSORB>
__global__ void __launch_bounds__(kBlock, STB_RK_MINB) rk_kernel(const Params p)
{
    extern __shared__ double sh_hist[];
    const bool hist = ABSORB && p.hist != nullptr;
    if (hist) {
        for (int b = threadIdx.x; b < p.nbins; b += kBlock) sh_hist[b] = 0.0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const Medium &m = p.m;
    const double umax = (double)(m.nx - 1), vmax = (double)(m.ny - 1);

    long long ray = -1;
    double x = 0, y = 0, px = 0, py = 0, phi = 0, sig = 0, tau = 0, h = 0, xs = 1, ys = 1, amp2 = 1, P = 1, n2min = 0;
    double f1x = 0, f1y = 0, a1 = 0;
    int nacc = 0, nrej = 0, status = 0;
    bool after_reject = false;
    // landing on the stop line: the step that crosses it is repeated with shorter lengths
    // (Illinois regula falsi on the step length) until its end state lies on the line
    // (the bracket lives in shared memory: it is touched on a handful of attempts per ray and
    // would otherwise cost ten registers in the hot loop)
    __shared__ double s_land[5 * kBlock];
    double &hland = s_land[threadIdx.x], &hl = s_land[kBlock + threadIdx.x], &dl = s_land[2 * kBlock + threadIdx.x];
    double &hu = s_land[3 * kBlock + threadIdx.x], &du = s_land[4 * kBlock + threadIdx.x];
    int land = 0, side = 0;
    int turn = -1;

    auto inside = [&](double xx, double yy) -> bool {
        const double u = (xx - m.xorg) * m.rdx, v = (yy - m.yorg) * m.rdy;
        return u >= 0.0 && u < umax && v >= 0.0 && v < vmax;
    };
    auto store_state = [&]() {
        if (p.traj && nacc < p.L) {
            double *tp = p.traj + (ray * p.L + nacc) * 4;
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(tp), "d"(x), "d"(y), "d"(px), "d"(py) : "memory");
            if (p.traj_aux) {
                double *ap = p.traj_aux + (ray * p.L + nacc) * 2;
                asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(ap), "d"(tau), "d"(sig) : "memory");
            }
        }
    };
    auto retire = [&]() {
        const long long n = p.nrays;
        if (p.end_state) {
            p.end_state[ray] = x; p.end_state[n + ray] = y; p.end_state[2 * n + ray] = px; p.end_state[3 * n + ray] = py;
        }
        if (p.tau_end) p.tau_end[ray] = tau;
        if (p.phi_end) p.phi_end[ray] = phi;
        if (p.sig_end) p.sig_end[ray] = sig;
        if (p.n2_min) p.n2_min[ray] = n2min;
        if (p.ham_end) {
            double fx, fy, al, n2;
            rk_eval<INTERP, MODE, false>(m, xs, ys, x, y, fx, fy, al, n2);
            p.ham_end[ray] = fma(px, px, py * py) - n2;
        }
        if (p.turn_index) p.turn_index[ray] = turn;
        if (p.turn_state && turn < 0) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            p.turn_state[ray] = qnan; p.turn_state[n + ray] = qnan; p.turn_state[2 * n + ray] = qnan; p.turn_state[3 * n + ray] = qnan;
        }
        if (p.nacc) p.nacc[ray] = nacc;
        if (p.nrej) p.nrej[ray] = nrej;
        if (p.status) p.status[ray] = status;
    };
    // returns false when the ray is over before its first step
    auto launch = [&](long long r) -> bool {
        ray = r;
        x = p.x0[r]; y = p.y0[r]; px = p.px0[r]; py = p.py0[r];
        xs = p.xscale ? p.xscale[r] : 1.0;
        ys = p.yscale ? p.yscale[r] : 1.0;
        const double a = p.amp ? p.amp[r] : 1.0;
        amp2 = a * a; P = amp2;
        phi = 0.0; sig = 0.0; tau = 0.0; h = p.h0;
        nacc = 0; nrej = 0; after_reject = false; land = 0; turn = -1;
        double n2;
        rk_eval<INTERP, MODE, ABSORB>(m, xs, ys, x, y, f1x, f1y, a1, n2);
        n2min = n2;
        store_state();
        if (!inside(x, y)) { status = START_OUTSIDE; retire(); return false; }
        if (p.max_steps <= 0) { status = MAX_STEPS; retire(); return false; }
        return true;
    };

    // one step attempt; false when the ray ended (outputs written)
    auto attempt = [&]() -> bool {
        double hh = land ? hland : h;
        bool last = false;
        if (tau + hh >= p.tau_max) { hh = p.tau_max - tau; last = true; }
        // stage derivatives: position <- kp (the stage's p), momentum <- kf.  The absorption coefficient
        // of a stage enters only the optical-depth quadrature, which is accumulated as the stages come
        // (same order as a sum over stored stages) instead of keeping seven more values in registers.
        double kpx[7], kpy[7], kfx[7], kfy[7];
        kpx[0] = px; kpy[0] = py; kfx[0] = f1x; kfy[0] = f1y;
        double ssig = cA[5][0] * a1, alast = a1;
        double xn = x, yn = y, pxn = px, pyn = py, n2n = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double sx = cA[i][0] * kpx[0], sy = cA[i][0] * kpy[0], sfx = cA[i][0] * kfx[0], sfy = cA[i][0] * kfy[0];
#pragma unroll
            for (int j = 1; j <= i; ++j) {
                sx = fma(cA[i][j], kpx[j], sx);
                sy = fma(cA[i][j], kpy[j], sy);
                sfx = fma(cA[i][j], kfx[j], sfx);
                sfy = fma(cA[i][j], kfy[j], sfy);
            }
            xn = fma(hh, sx, x); yn = fma(hh, sy, y);
            pxn = fma(hh, sfx, px); pyn = fma(hh, sfy, py);
            kpx[i + 1] = pxn; kpy[i + 1] = pyn;
            rk_eval<INTERP, MODE, ABSORB>(m, xs, ys, xn, yn, kfx[i + 1], kfy[i + 1], alast, n2n);
            if (ABSORB && i < 5) ssig = fma(cA[5][i + 1], alast, ssig);
        }
        // (xn, yn, pxn, pyn) is the 5th-order solution (row 5 = b); stage 7 is its derivative
        double ex = cE[0] * kpx[0], ey = cE[0] * kpy[0], efx = cE[0] * kfx[0], efy = cE[0] * kfy[0];
#pragma unroll
        for (int j = 1; j < 7; ++j) {
            ex = fma(cE[j], kpx[j], ex);
            ey = fma(cE[j], kpy[j], ey);
            efx = fma(cE[j], kfx[j], efx);
            efy = fma(cE[j], kfy[j], efy);
        }
        // err = max_i |e_i| / sc_i, sc_i = atol + rtol max(|y_i|, |ynew_i|): the largest ratio is found
        // by cross-multiplication (the scales are positive), so the step costs one division, not four
        const double ah = fabs(hh);
        const double scx = fma(p.rtol, max2(fabs(x), fabs(xn)), p.atol_r);
        const double scy = fma(p.rtol, max2(fabs(y), fabs(yn)), p.atol_r);
        const double scpx = fma(p.rtol, max2(fabs(px), fabs(pxn)), p.atol_p);
        const double scpy = fma(p.rtol, max2(fabs(py), fabs(pyn)), p.atol_p);
        const double q0 = ah * fabs(ex), q1 = ah * fabs(ey), q2 = ah * fabs(efx), q3 = ah * fabs(efy);
        const double esum = (q0 + q1) + (q2 + q3);   // a NaN anywhere must reject the step
        double en = q0, ed = scx;
        if (q1 * ed > en * scy) { en = q1; ed = scy; }
        if (q2 * ed > en * scpx) { en = q2; ed = scpx; }
        if (q3 * ed > en * scpy) { en = q3; ed = scpy; }
        const double err = (esum == esum) ? en / ed : esum;
        if (!(err <= 1.0)) {   // rejected (NaN included)
            ++nrej;
            const double fac = (err == err) ? max2(0.2, 0.9 * inv_root5(err)) : 0.2;
            h = hh * fac;
            after_reject = true;
            land = 0;              // a landing trial that fails the error test: resume stepping, re-detect
            if (!(h >= p.hmin)) { status = STEP_UNDERFLOW; retire(); return false; }
            return true;
        }
        bool on_line = false;
        if (p.stop_axis) {
            const double d0 = (p.stop_axis == 1 ? y : x) - p.stop_val;
            const double d1 = (p.stop_axis == 1 ? yn : xn) - p.stop_val;
            if (!land) {
                if (d0 != 0.0 && ((d0 < 0.0) != (d1 < 0.0) || d1 == 0.0)) {      // the line is crossed in this step
                    if (fabs(d1) <= p.stop_tol) {
                        on_line = true;
                    } else {
                        land = 1; side = 0;
                        hl = 0.0; dl = d0; hu = hh; du = d1;
                        hland = hh * (d0 / (d0 - d1));
                        ++nrej;
                        return true;
                    }
                }
            } else if (fabs(d1) <= p.stop_tol || land >= 12) {
                on_line = true;
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
                return true;
            }
        }
        // accepted: quadratures with the 5th-order weights
        double sphi = cA[5][0] * fma(kpx[0], kpx[0], kpy[0] * kpy[0]);
#pragma unroll
        for (int j = 1; j < 6; ++j) sphi = fma(cA[5][j], fma(kpx[j], kpx[j], kpy[j] * kpy[j]), sphi);
        const double xm = 0.5 * (x + xn), ym = 0.5 * (y + yn);
        // turning point (the parity kernel's definition, sidtrace_b200.h turn_index): first accepted state at
        // which p_y has changed sign; written straight to memory, it costs no register in the loop
        if (turn < 0 && ((py > 0.0 && pyn <= 0.0) || (py < 0.0 && pyn >= 0.0))) {
            turn = nacc + 1;
            if (p.turn_state) {
                const long long nn = p.nrays;
                p.turn_state[ray] = xn; p.turn_state[nn + ray] = yn; p.turn_state[2 * nn + ray] = pxn; p.turn_state[3 * nn + ray] = pyn;
            }
        }
        x = xn; y = yn; px = pxn; py = pyn;
        phi = fma(hh, sphi, phi);
        sig = fma(hh, ssig, sig);
        tau = last ? p.tau_max : tau + hh;
        f1x = kfx[6]; f1y = kfy[6]; a1 = alast;
        n2min = min2(n2n, n2min);
        ++nacc;
        store_state();
        if (hist) {
            const double Pn = amp2 * exp(-sig);
            const double dP = P - Pn;
            P = Pn;
            const double rx = xm - p.xaxis, ry = ym - p.yaxis;
            const double rho = sqrt(fma(rx, rx, ry * ry)) * p.r_rlcs;
            const double fb = rho * p.bins_per_rho;
            const int b = (fb < (double)p.nbins) ? (int)fb : -1;
            rk_hist_add(sh_hist, b, dP);
        }
        // no growth right after a rejection: keeps the controller from bouncing at the kinks of the
        // interpolant (cell edges)
        const double fmx = after_reject ? 1.0 : 5.0;
        const double fac = min2(0.9 * inv_root5(err), fmx);
        after_reject = false;
        h = min2(hh * fac, p.hmax);
        if (on_line) { status = STOP_LINE; retire(); return false; }
        if (!inside(x, y)) { status = LEFT_DOMAIN; retire(); return false; }
        if (last) { status = TAU_MAX; retire(); return false; }
        if (nacc >= p.max_steps) { status = MAX_STEPS; retire(); return false; }
        return true;
    };

    // Persistent warps: vote, claim rays for the idle lanes with one aggregated atomic once refill_min lanes are
    // idle, step.  refill_min = 32 ("warp tiles"): the warp owns 32 CONSECUTIVE rays at a time; neighbouring rays
    // of a beam sit in the same or adjacent cells, so each stencil load of the warp touches 4 sectors instead of 16
    // (the L1 data pipe is the co-limiter of this kernel) -- at the price of idle lanes while the slowest ray of the
    // tile finishes.  refill_min = 1: lane refill.  In between (default 24): most of the locality, and a straggler
    // never holds more than refill_min - 1 lanes idle.
    bool active = false, more = true;
    for (;;) {
        unsigned act = __ballot_sync(FULL, active);
        if (more && __popc(~act) >= p.refill_min) {
            const unsigned idle = ~act;
            const int leader = __ffs(idle) - 1;
            const int want = __popc(idle);
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(p.counter, (unsigned long long)want);
            base = __shfl_sync(FULL, base, leader);
            if (!active) {
                const long long r = (long long)base + __popc(idle & ((1u << lane) - 1u));
                if (r < p.nrays) active = launch(r);
            }
            more = (long long)(base + want) < p.nrays;
            act = __ballot_sync(FULL, active);
            if (act == 0 && more) continue;   // every claimed ray ended at launch: claim again
        }
        if (act == 0) break;
        if (active) {
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                if (!attempt()) { active = false; break; }
            }
        }
    }

    if (hist) {
        __syncthreads();
        for (int b = threadIdx.x; b < p.nbins; b += kBlock) {
            const double v = sh_hist[b];
            if (v != 0.0) atomicAdd(&p.hist[b], v);
        }
    }
}

__global__ void interleave2_kernel(const double *__restrict__ a, const double *__restrict__ b, double2 *__restrict__ out,
                                   long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_double2(a[i], b[i]);
}

// Y^2 = (omega_ce/omega)^2 on the nodes of the medium's grid from a tabulated equilibrium (include/sidtrace_b200_rk.h,
// stb_equilibrium_t; the table of the reference's Equilibrium class, equilibrium.py:11-49, interpolated bilinearly
// as its RectBivariateSpline(kx=1, ky=1) does).  One thread per node; plain IEEE operations in a fixed order (the
// library is compiled -fmad=false), so oracle/rk_oracle.c reproduces the table bit for bit.
struct EqTable {
    const double *psi;     // [nr][nz]
    const double *fpol;    // [nf]
    int nr, nz, nf;
    double rleft, zlow, rdr, rdz;     // first node, 1/spacing (1/m)
    double simagx, rdpsi;             // F grid: psi = simagx + k / rdpsi
    double r0_cm, z0_cm, ypt;
};
__global__ void equilibrium_y2_kernel(const EqTable e, int nx, int ny, double xorg, double yorg, double dx, double dy,
                                      double *__restrict__ y2)
{
    const long long n = (long long)nx * ny;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(k / nx), i = (int)(k - (long long)j * nx);
        const double R = (e.r0_cm + (xorg + dx * (double)i)) * 0.01;
        const double Z = ((yorg + dy * (double)j) - e.z0_cm) * 0.01;
        const double u = (R - e.rleft) * e.rdr, v = (Z - e.zlow) * e.rdz;
        // cell clamped to the table, fraction not: outside, the edge cell's bilinear patch is continued
        const int iu = max(0, min(__double2int_rd(u), e.nr - 2));
        const int iv = max(0, min(__double2int_rd(v), e.nz - 2));
        const double t = u - (double)iu, s = v - (double)iv;
        const double *q = e.psi + (long long)iu * e.nz + iv;
        const double p00 = q[0], p01 = q[1], p10 = q[e.nz], p11 = q[e.nz + 1];    // p[R index][Z index]
        const double lo = p00 + t * (p10 - p00), hi = p01 + t * (p11 - p01);
        const double psi = lo + s * (hi - lo);                                    // bilinear, as equilibrium.py:46-49
        // grad psi: central differences at the four corner nodes (one-sided on the border), interpolated bilinearly --
        // a CONTINUOUS poloidal field.  (The gradient of the bilinear patch itself jumps at every cell edge of the psi
        // grid; sampled onto the medium's finer grid those jumps cost the step controller 7x the steps.)
        double gR[2][2], gZ[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ia = iu + a, ib = iv + b;
                const int a0 = max(ia - 1, 0), a1 = min(ia + 1, e.nr - 1), b0 = max(ib - 1, 0), b1 = min(ib + 1, e.nz - 1);
                gR[a][b] = (e.psi[(long long)a1 * e.nz + ib] - e.psi[(long long)a0 * e.nz + ib]) * (a1 - a0 == 2 ? 0.5 * e.rdr : e.rdr);
                gZ[a][b] = (e.psi[(long long)ia * e.nz + b1] - e.psi[(long long)ia * e.nz + b0]) * (b1 - b0 == 2 ? 0.5 * e.rdz : e.rdz);
            }
        const double rlo = gR[0][0] + t * (gR[1][0] - gR[0][0]), rhi = gR[0][1] + t * (gR[1][1] - gR[0][1]);
        const double psiR = rlo + s * (rhi - rlo);            // dpsi/dR
        const double zlo = gZ[0][0] + t * (gZ[1][0] - gZ[0][0]), zhi = gZ[0][1] + t * (gZ[1][1] - gZ[0][1]);
        const double psiZ = zlo + s * (zhi - zlo);            // dpsi/dZ
        // F(psi): linear on the uniform psi grid, held at the ends
        double w = (psi - e.simagx) * e.rdpsi;
        w = w > 0.0 ? w : 0.0;
        const double wmax = (double)(e.nf - 1);
        w = w < wmax ? w : wmax;
        const int iw = min(__double2int_rd(w), e.nf - 2);
        const double fw = w - (double)iw;
        const double F = e.fpol[iw] + fw * (e.fpol[iw + 1] - e.fpol[iw]);
        const double rR = 1.0 / R;
        const double BR = -(psiZ * rR), BZ = psiR * rR, Bp = F * rR;
        const double B2 = (BR * BR + BZ * BZ) + Bp * Bp;
        y2[k] = (e.ypt * e.ypt) * B2;
    }
}

// Cubic B-spline prefilter along one axis: for every line, solve c[i-1] + 4 c[i] + c[i+1] = 6 f[i]
// with the natural end conditions c[0] = f[0], c[n-1] = f[n-1] and extend by c[-1] = 2c[0] - c[1],
// c[n] = 2c[n-1] - c[n-2].  One thread per line (Thomas algorithm; R[i] are the data-independent
// pivots' reciprocals, R[1] = 1/4, R[i] = 1/(4 - R[i-1])).  Output index 0 holds c[-1].
__global__ void bspline_prefilter_kernel(const double *__restrict__ in, long long in_line, long long in_elem,
                                         double *__restrict__ out, long long out_line, long long out_elem,
                                         const double *__restrict__ R, int n, int nlines)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const double *f = in + (long long)l * in_line;
    double *c = out + (long long)l * out_line + out_elem;   // c[i] at c + i*out_elem, i = -1 .. n
    const double f0 = f[0], fl = f[(long long)(n - 1) * in_elem];
    c[0] = f0;
    c[(long long)(n - 1) * out_elem] = fl;
    if (n > 2) {
        double dprev = 0.0;
        for (int i = 1; i <= n - 2; ++i) {
            double rhs = 6.0 * f[(long long)i * in_elem];
            if (i == 1) rhs -= f0;
            if (i == n - 2) rhs -= fl;
            const double d = (rhs - dprev) * R[i];
            c[(long long)i * out_elem] = d;
            dprev = d;
        }
        double cnext = c[(long long)(n - 2) * out_elem];
        for (int i = n - 3; i >= 1; --i) {
            const double ci = c[(long long)i * out_elem] - R[i] * cnext;
            c[(long long)i * out_elem] = ci;
            cnext = ci;
        }
    }
    c[-out_elem] = 2.0 * c[0] - c[out_elem];
    c[(long long)n * out_elem] = 2.0 * c[(long long)(n - 1) * out_elem] - c[(long long)(n - 2) * out_elem];
}

}  // namespace stbrk
