// field_b200.cu -- the O(rays) tail of the beam-level consumers on the device:
// Beam.field_r (beam.py:347-444) after the per-ray crossings are known, and field_result
// (beam.py:447-472).  Declared in include/sidtrace_b200.h; linked into libsidtrace_b200.so.
//
//   stb_field_pages_links   host, O(n log P): the "pages" of beam.py:374-398 as per-ray links
//   stb_tube_field          ray-tube amplitude sqrt(d0/d1) and complex field per ray (beam.py:404-427)
//   stb_hole_fill           rays of dropped pages get their neighbours' amplitude (beam.py:436-440)
//   stb_overlap             TX field x analytic RX pattern, |f|-weighted centre, signal = sum f dx (beam.py:448-461)
//   stb_gauss_fit_sums      normal-equation sums of the Gaussian fit that replaces curve_fit (beam.py:463-467)
//
// Tolerance-level parity (1e-9): the reference computes these with NumPy / SciPy (libm, pairwise sums,
// MINPACK); nothing here feeds back into the tracer.
#include "../../include/sidtrace_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

extern "C" int stb_internal_fail(int code, const char *msg);
extern "C" void stb_internal_count_launch(void);

namespace {

int cfail(const char *what, cudaError_t e)
{
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    return stb_internal_fail(STB_ERR_CUDA, buf);
}

int grid_for(long long n, int threads)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long b = (n + threads - 1) / threads;
    if (b > (long long)sms * 16) b = (long long)sms * 16;
    return b < 1 ? 1 : (int)b;
}

__global__ void tube_kernel(const stb_tube_t a)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long n = a.n;
    for (; i < n; i += step) {
        const double amp0 = a.amp0[i];
        double amp1 = amp0, fre = 0.0, fim = 0.0;
        if (a.page_size[i] >= 2) {                       // single-ray pages are dropped (beam.py:392-398)
            const int l = a.pred[i], r = a.succ[i];
            const double x0 = a.x0[i], y0 = a.y0[i], xc = a.xc[i], yc = a.yc[i];
            double d0l = 0, d0r = 0, d1l = 0, d1r = 0;
            if (l >= 0) {
                d0l = hypot(x0 - a.x0[l], y0 - a.y0[l]);
                // the left neighbour sampled at MY phase: its A target (its right-hand neighbour is me)
                d1l = hypot(a.lookup_xy[0 * n + l] - xc, a.lookup_xy[1 * n + l] - yc);
            }
            if (r >= 0) {
                d0r = hypot(x0 - a.x0[r], y0 - a.y0[r]);
                d1r = hypot(a.lookup_xy[2 * n + r] - xc, a.lookup_xy[3 * n + r] - yc);
            }
            // page ends: the missing neighbour's distance is the other one's (beam.py:419-420)
            if (l < 0) { d0l = d0r; d1l = d1r; }
            if (r < 0) { d0r = d0l; d1r = d1l; }
            double a1 = sqrt((d0l + d0r) / (d1l + d1r));   // ray-tube divergence: 2-D, 1/sqrt(separation)
            if (a1 != a1) a1 = amp0;
            amp1 = a1;
            double sn, cs;
            sincos(a.ph[i], &sn, &cs);
            fre = amp0 * a1 * cs;                           // amp0 * amp1 * exp(-1j * ph)
            fim = -(amp0 * a1 * sn);
        }
        a.amp1[i] = amp1;
        a.field[2 * i] = fre;
        a.field[2 * i + 1] = fim;
    }
}

__global__ void hole_fill_kernel(long long n, const double *__restrict__ amp0, const double *__restrict__ ph,
                                 const double *__restrict__ amp1_in, double *__restrict__ amp1_out, double *__restrict__ field)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        double a1 = amp1_in[i];
        if (field[2 * i] == 0.0 && field[2 * i + 1] == 0.0) {
            const long long nx = i + 1 < n ? i + 1 : 0, pv = i > 0 ? i - 1 : n - 1;     // np.roll wraps around
            a1 = (amp1_in[nx] + amp1_in[pv]) / 2;
            if (a1 != a1) a1 = amp0[i];
            double sn, cs;
            sincos(ph[i], &sn, &cs);
            field[2 * i] = amp0[i] * a1 * cs;
            field[2 * i + 1] = -(amp0[i] * a1 * sn);
        } else if (a1 != a1) {
            a1 = amp0[i];
        }
        amp1_out[i] = a1;
    }
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    if (w == 0) for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;     // valid in thread 0
}

__global__ void overlap_kernel(const stb_overlap_t a)
{
    __shared__ double sh[32];
    const long long n = a.n;
    double s_abs = 0, s_absx = 0, s_re = 0, s_im = 0, mx = 0;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const double pi = 3.14159265358979323846;
    for (; i < n; i += step) {
        const double x = a.x[i];
        // analytic RX pattern at (x, y_rx): beam.py:103-131 (lab2beam), 176-206 (_field), n = 1
        const double dxl = x - a.bx0, dyl = a.y_rx - a.by0;
        const double xx = a.bky * dxl - a.bkx * dyl;
        const double yy = a.bkx * dxl + a.bky * dyl - a.focus;
        const double y_r = a.rayleigh;
        const double wa = a.waist * sqrt(1.0 + yy * yy / (y_r * y_r));
        const double inv_r = yy != 0.0 ? 1.0 / (yy * (1.0 + y_r * y_r / (yy * yy))) : 0.0;      // r = inf on the waist
        const double psi = atan(yy / y_r) / 2;
        const double amp = sqrt(a.waist / wa) * exp(-xx * xx / (wa * wa));
        const double phase = 2 * pi * yy / a.wave_len + pi * xx * xx * inv_r / a.wave_len - psi * a.gouy;
        double sn, cs;
        sincos(phase, &sn, &cs);
        const double rre = amp * cs, rim = -(amp * sn);
        const double tre = a.field_tx[2 * i], tim = a.field_tx[2 * i + 1];
        const double fre = tre * rre - tim * rim, fim = tre * rim + tim * rre;
        a.field[2 * i] = fre;
        a.field[2 * i + 1] = fim;
        const double af = hypot(fre, fim);
        a.absf[i] = af;
        // dist = ((roll(x,-1) - x) + (x - roll(x,1))) / 2 with dist[0] = dist[1], dist[-1] = dist[-2] (beam.py:455-457)
        long long j = i;
        if (n >= 3) { if (i == 0) j = 1; else if (i == n - 1) j = n - 2; }
        const long long jn = j + 1 < n ? j + 1 : 0, jp = j > 0 ? j - 1 : n - 1;
        const double dist = ((a.x[jn] - a.x[j]) + (a.x[j] - a.x[jp])) / 2;
        s_abs += af; s_absx += af * x; s_re += fre * dist; s_im += fim * dist;
        mx = fmax(mx, af);
    }
    double v;
    v = block_sum(s_abs, sh);  if (threadIdx.x == 0) atomicAdd(&a.sums[0], v);
    v = block_sum(s_absx, sh); if (threadIdx.x == 0) atomicAdd(&a.sums[1], v);
    v = block_sum(s_re, sh);   if (threadIdx.x == 0) atomicAdd(&a.sums[2], v);
    v = block_sum(s_im, sh);   if (threadIdx.x == 0) atomicAdd(&a.sums[3], v);
    // max of non-negative doubles = max of their bit patterns
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)&a.sums[4], (unsigned long long)__double_as_longlong(mx));
}

// h0, h1, h2 != 0: forward-difference Jacobian columns (model(p + h_j e_j) - model(p)) / h_j, the way
// MINPACK's fdjac2 forms them (curve_fit without jac=); 0: analytic columns.
__global__ void gauss_sums_kernel(const double *__restrict__ x, const double *__restrict__ y, long long n,
                                  double amp, double xc, double sigma, double h0, double h1, double h2, double *out)
{
    __shared__ double sh[32];
    double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    const bool fd = h0 != 0.0;
    for (; i < n; i += step) {
        const double xi = x[i];
        const double d = xi - xc;
        // amp * np.exp(-(xa - xc) ** 2 / 2 / sigma ** 2), operation for operation
        const double g = amp * exp(-(d * d) / 2 / (sigma * sigma));
        const double r = g - y[i];
        double ja, jb, jc;
        if (fd) {
            ja = ((amp + h0) * exp(-(d * d) / 2 / (sigma * sigma)) - g) / h0;
            const double d1 = xi - (xc + h1);
            jb = (amp * exp(-(d1 * d1) / 2 / (sigma * sigma)) - g) / h1;
            const double s2 = sigma + h2;
            jc = (amp * exp(-(d * d) / 2 / (s2 * s2)) - g) / h2;
        } else {
            const double is2 = 1.0 / (sigma * sigma);
            ja = g / amp; jb = g * d * is2; jc = g * d * d * is2 / sigma;
        }
        s[0] += ja * ja; s[1] += ja * jb; s[2] += ja * jc; s[3] += jb * jb; s[4] += jb * jc; s[5] += jc * jc;
        s[6] += ja * r; s[7] += jb * r; s[8] += jc * r; s[9] += r * r;
    }
    for (int k = 0; k < 10; ++k) {
        const double v = block_sum(s[k], sh);
        if (threadIdx.x == 0) atomicAdd(&out[k], v);
    }
}

}  // namespace

extern "C" {

int stb_field_pages_links(const double *xout, int64_t n, int32_t *pred, int32_t *succ, int32_t *page_size, int64_t *npages)
{
    // beam.py:374-398 walks the rays once per page: a ray joins the page being walked when it lies
    // beyond the page's last member and belongs to no page yet; the first ray that does not seeds the
    // next page.  That is first-fit: every ray goes to the FIRST page whose last member it lies beyond.
    // The last members are non-increasing in the page index (a ray reaches page k+1 only after page k
    // refused it), so the page is found by bisection: O(n log P) instead of O(n P).
    if (!xout || !pred || !succ || !page_size || !npages || n < 0 || n > 0x7fffffffLL)
        return stb_internal_fail(STB_ERR_ARG, "stb_field_pages_links: bad argument");
    std::vector<double> top;
    std::vector<int32_t> last, count, page((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        const double v = xout[j];
        size_t lo = 0, hi = top.size();          // first k with top[k] < v
        while (lo < hi) {
            const size_t mid = (lo + hi) >> 1;
            if (top[mid] < v) hi = mid; else lo = mid + 1;
        }
        if (lo == top.size()) {
            top.push_back(v); last.push_back((int32_t)j); count.push_back(1);
            pred[j] = -1;
        } else {
            pred[j] = last[lo];
            succ[last[lo]] = (int32_t)j;
            top[lo] = v; last[lo] = (int32_t)j; ++count[lo];
        }
        succ[j] = -1;
        page[(size_t)j] = (int32_t)lo;
    }
    int64_t kept = 0;
    for (int32_t c : count) kept += c >= 2;
    for (int64_t j = 0; j < n; ++j) page_size[j] = count[(size_t)page[(size_t)j]];
    *npages = kept;
    return STB_OK;
}

int stb_tube_field(const stb_tube_t *a, void *stream)
{
    if (!a || a->struct_size != sizeof(stb_tube_t)) return stb_internal_fail(STB_ERR_ARG, "stb_tube_field: bad struct");
    const void *ptrs[] = {a->pred, a->succ, a->page_size, a->x0, a->y0, a->xc, a->yc, a->ph, a->amp0, a->lookup_xy, a->amp1, a->field};
    for (const void *q : ptrs) if (!q && a->n > 0) return stb_internal_fail(STB_ERR_ARG, "stb_tube_field: NULL pointer");
    if (a->n <= 0) return STB_OK;
    tube_kernel<<<grid_for(a->n, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    stb_internal_count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? STB_OK : cfail("stb_tube_field", e);
}

int stb_hole_fill(int64_t n, const double *amp0, const double *ph, const double *amp1_in, double *amp1_out, double *field, void *stream)
{
    if (n > 0 && (!amp0 || !ph || !amp1_in || !amp1_out || !field || amp1_in == amp1_out))
        return stb_internal_fail(STB_ERR_ARG, "stb_hole_fill: bad argument");
    if (n <= 0) return STB_OK;
    hole_fill_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, amp0, ph, amp1_in, amp1_out, field);
    stb_internal_count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? STB_OK : cfail("stb_hole_fill", e);
}

int stb_overlap(const stb_overlap_t *a, void *stream)
{
    if (!a || a->struct_size != sizeof(stb_overlap_t)) return stb_internal_fail(STB_ERR_ARG, "stb_overlap: bad struct");
    if (a->n > 0 && (!a->x || !a->field_tx || !a->field || !a->absf || !a->sums))
        return stb_internal_fail(STB_ERR_ARG, "stb_overlap: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(a->sums, 0, 8 * sizeof(double), s);
    if (e != cudaSuccess) return cfail("stb_overlap", e);
    if (a->n <= 0) return STB_OK;
    overlap_kernel<<<grid_for(a->n, 256), 256, 0, s>>>(*a);
    stb_internal_count_launch();
    e = cudaGetLastError();
    return e == cudaSuccess ? STB_OK : cfail("stb_overlap", e);
}

int stb_gauss_fit_sums(const double *x, const double *y, int64_t n, double amp, double xc, double sigma, const double *fd_step,
                       double *out, void *stream)
{
    if (!out || (n > 0 && (!x || !y))) return stb_internal_fail(STB_ERR_ARG, "stb_gauss_fit_sums: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, 10 * sizeof(double), s);
    if (e != cudaSuccess) return cfail("stb_gauss_fit_sums", e);
    if (n <= 0) return STB_OK;
    const double h0 = fd_step ? fd_step[0] : 0.0, h1 = fd_step ? fd_step[1] : 0.0, h2 = fd_step ? fd_step[2] : 0.0;
    if (fd_step && (h0 == 0.0 || h1 == 0.0 || h2 == 0.0)) return stb_internal_fail(STB_ERR_ARG, "stb_gauss_fit_sums: zero finite-difference step");
    gauss_sums_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, y, n, amp, xc, sigma, h0, h1, h2, out);
    stb_internal_count_launch();
    e = cudaGetLastError();
    return e == cudaSuccess ? STB_OK : cfail("stb_gauss_fit_sums", e);
}

}  // extern "C"
