// launch_kernels.cuh -- launch-state preparation on the device: what the
// reference's Ray.__init__ does per ray before calling trace (ray.py:34-68), for a whole beam.
//
//   k  = (kx0, ky0) / |k0| * w / c                       ray.py:38-45   normalise to the vacuum wave number
//   N  = sqrt(1 - n(x0, y0) / n_c)                        ray.py:62, den.py:76-82 with the cubic spline of n
//   k *= N                                               ray.py:63-64
//   xg = x2grid(x0), yg = y2grid(y0)                     ray.py:67-68, den.py:69-70 (piecewise-linear
//                                                        interp1d with extrapolation)
//
// The grid coordinates decide which cell a ray starts in, so x2grid/y2grid follow scipy's interp1d
// arithmetic operation for operation (bit-identical to the host path).  n(x0, y0) evaluates the
// bicubic B-spline that SciPy's RectBivariateSpline fitted on the host (FITPACK fpbisp/fpbspl
// restated); it agrees with SciPy's own evaluation to rounding (tolerance-level path, like the
// reference's interp2d it replaces).
#pragma once
#include <cuda_runtime.h>

namespace stb {

struct LaunchParams {
    long long n;
    const double *x0p, *y0p, *kxd, *kyd;          // physical start points, direction (any length)
    const double *xn, *yn;                          // grid node coordinates
    int nx, ny;
    const double *tx, *ty, *coef;                   // spline knots and coefficients (degree 3 x 3)
    int ntx, nty;
    double w, c, nc;
    double *xg, *yg, *kx, *ky;                      // outputs
};

// interp1d(x, arange(n), fill_value='extrapolate')(v) as SciPy >= 1.12 evaluates it
// (interp1d._call_linear): node hi = searchsorted(x, v) clipped to [1, n-1], lo = hi - 1,
//   ((v - x[lo]) / (x[hi] - x[lo])) * hi  +  ((x[hi] - v) / (x[hi] - x[lo])) * lo.
__device__ __forceinline__ double to_grid(const double *__restrict__ x, int n, double v)
{
    int lo = 0, hi = n;                             // first index with x[idx] >= v  (searchsorted 'left')
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] < v) lo = mid + 1; else hi = mid;
    }
    int b = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
    const int a = b - 1;
    const double w = __dsub_rn(x[b], x[a]);
    const double up = __dmul_rn(__ddiv_rn(__dsub_rn(v, x[a]), w), (double)b);
    const double dn = __dmul_rn(__ddiv_rn(__dsub_rn(x[b], v), w), (double)a);
    return __dadd_rn(up, dn);
}

// FITPACK fpbspl for degree 3: the four non-zero B-splines at x, knot interval l (t[l] <= x < t[l+1]).
__device__ __forceinline__ void bspl3(const double *__restrict__ t, double x, int l, double h[4])
{
    double hh[3];
    h[0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 0; i < j; ++i) {
            const int li = l + i + 1, lj = li - j;
            if (t[li] != t[lj]) {
                const double f = hh[i] / (t[li] - t[lj]);
                h[i] = h[i] + f * (t[li] - x);
                h[i + 1] = f * (x - t[lj]);
            } else {
                h[i + 1] = 0.0;
            }
        }
    }
}

__device__ __forceinline__ int knot_interval(const double *__restrict__ t, int nt, double x)
{
    // largest l in [3, nt-5] with t[l] <= x  (FITPACK fpbisp; x already clamped to [t[3], t[nt-4]])
    int lo = 3, hi = nt - 5;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void launch_states_kernel(const LaunchParams p)
{
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; r < p.n; r += step) {
        const double x0 = p.x0p[r], y0 = p.y0p[r];
        // density at the start point: clamp to the spline's domain like fpbisp does
        double xs = x0, ys = y0;
        const double xb = p.tx[3], xe = p.tx[p.ntx - 4], yb = p.ty[3], ye = p.ty[p.nty - 4];
        xs = xs < xb ? xb : (xs > xe ? xe : xs);
        ys = ys < yb ? yb : (ys > ye ? ye : ys);
        const int lx = knot_interval(p.tx, p.ntx, xs), ly = knot_interval(p.ty, p.nty, ys);
        double hx[4], hy[4];
        bspl3(p.tx, xs, lx, hx);
        bspl3(p.ty, ys, ly, hy);
        const int ncy = p.nty - 4;
        double den = 0.0;
        for (int i = 0; i < 4; ++i) {
            const double *row = p.coef + (long long)(lx - 3 + i) * ncy + (ly - 3);
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s += row[j] * hy[j];
            den += hx[i] * s;
        }
        // den.py:71: interp2d(..., fill_value=1e12) -- outside the grid the reference reads 1e12
        if (x0 < xb || x0 > xe || y0 < yb || y0 > ye) den = 1e12;
        const double refr = sqrt(1.0 - den / p.nc);
        // direction -> vacuum wave vector -> local wave vector (same operation order as launch_states)
        const double kxd = p.kxd[r], kyd = p.kyd[r];
        const double kn = sqrt(kxd * kxd + kyd * kyd);
        p.kx[r] = kxd / kn * p.w / p.c * refr;
        p.ky[r] = kyd / kn * p.w / p.c * refr;
        p.xg[r] = to_grid(p.xn, p.nx, x0);
        p.yg[r] = to_grid(p.yn, p.ny, y0);
    }
}

// ---------------------------------------------------------------------------------------------
// Beam.rays on the device (beam.py:293-331): start points equidistant in arc length along the
// wavefront, unit normals from the reference's fine-grid central differences, Gaussian amplitudes.
// Everything is a function of the ray index and a few 1000-node tables, so 10^8 rays cost
// milliseconds instead of minutes of NumPy.
struct BeamLaunchParams {
    long long n;            // rnum
    long long m_fine;       // lcm(rnum, nfront)
    long long stride;       // m_fine / rnum
    const double *lax, *wfx;        // [nw] arc length at the wavefront nodes, their beam-frame x
    int nw;
    const double *t, *c;            // cubic B-spline of the wavefront y(x): knots [nt], coefficients [nt-4]
    int nt;
    double bx0, by0, bkx, bky, focus;     // beam frame (beam.py:103-131)
    double w0, y_r;                       // waist, Rayleigh length at the edge index (beam.py:176-188)
    double *x, *y, *kx, *ky, *amp;        // outputs (lab frame)
};

// np.interp / interp1d(kind='linear') inside the table range
__device__ __forceinline__ double lin_interp(const double *__restrict__ xp, const double *__restrict__ fp, int n, double v)
{
    int lo = 0, hi = n - 1;                         // largest j with xp[j] <= v, clipped to [0, n-2]
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (xp[mid] <= v) lo = mid; else hi = mid - 1;
    }
    if (lo > n - 2) return fp[n - 1];
    const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    return slope * (v - xp[lo]) + fp[lo];
}

__device__ __forceinline__ double spl1(const double *__restrict__ t, const double *__restrict__ c, int nt, double x)
{
    const double xb = t[3], xe = t[nt - 4];
    x = x < xb ? xb : (x > xe ? xe : x);
    const int l = knot_interval(t, nt, x);
    double h[4];
    bspl3(t, x, l, h);
    double s = 0.0;
    for (int i = 0; i < 4; ++i) s += c[l - 3 + i] * h[i];
    return s;
}

__global__ void beam_launch_kernel(const BeamLaunchParams p)
{
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const double ltot = p.lax[p.nw - 1];
    const double step_r = p.n > 1 ? ltot / (double)(p.n - 1) : 0.0;        // np.linspace(0, ltot, n)
    const double step_f = p.m_fine > 1 ? ltot / (double)(p.m_fine - 1) : 0.0;
    auto lab = [&](double xb, double yb, double &xl, double &yl) {            // beam2lab
        const double yf = yb + p.focus;
        xl = p.bky * xb + p.bkx * yf + p.bx0;
        yl = -p.bkx * xb + p.bky * yf + p.by0;
    };
    auto fine = [&](long long j, double &xl, double &yl) {
        const double tt = j == p.m_fine - 1 ? ltot : (double)j * step_f + 0.0;
        const double xb = lin_interp(p.lax, p.wfx, p.nw, tt);
        lab(xb, spl1(p.t, p.c, p.nt, xb), xl, yl);
    };
    for (; r < p.n; r += step) {
        const double tt = r == p.n - 1 ? ltot : (double)r * step_r;
        const double xb = lin_interp(p.lax, p.wfx, p.nw, tt);
        const double yb = spl1(p.t, p.c, p.nt, xb);
        double xl, yl;
        lab(xb, yb, xl, yl);
        const long long j = r * p.stride;
        const long long jlo = j > 0 ? j - 1 : 0, jhi = j < p.m_fine - 1 ? j + 1 : p.m_fine - 1;
        double xlo, ylo, xhi, yhi;
        fine(jlo, xlo, ylo);
        fine(jhi, xhi, yhi);
        const double width = (j == 0 || j == p.m_fine - 1) ? 1.0 : 2.0;       // np.gradient: one-sided at the ends
        double kx = -((yhi - ylo) / width), ky = (xhi - xlo) / width;
        const double nrm = sqrt(kx * kx + ky * ky);
        kx /= nrm; ky /= nrm;
        // Gaussian amplitude at the start point (beam.py:183-188), beam frame of the lab point
        const double xx = p.bky * (xl - p.bx0) - p.bkx * (yl - p.by0);
        const double yy = p.bkx * (xl - p.bx0) + p.bky * (yl - p.by0) - p.focus;
        const double wa = p.w0 * sqrt(1.0 + yy * yy / (p.y_r * p.y_r));
        p.x[r] = xl; p.y[r] = yl; p.kx[r] = kx; p.ky[r] = ky;
        p.amp[r] = sqrt(p.w0 / wa) * exp(-(xx * xx) / (wa * wa));
    }
}

}  // namespace stb
