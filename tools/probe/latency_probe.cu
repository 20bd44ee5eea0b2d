// latency_probe.cu -- dependent-issue latencies on the target GPU that bound the O(rays)-output trace kernels:
// one warp per SM, chains of dependent operations timed with clock64().  Standalone (nvcc only, no library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o latency_probe latency_probe.cu && ./latency_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int N = 4096;

__global__ void k_dadd(double *out, double a, long long *cyc)
{
    double v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = __dadd_rn(v, a);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dmul(double *out, double a, long long *cyc)
{
    double v = threadIdx.x + 1.0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = __dmul_rn(v, a);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma(double *out, double a, double b, long long *cyc)
{
    double v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = __fma_rn(v, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// F2I.F64.TRUNC -> I2F.F64 -> DADD round trip
__global__ void k_f2i(double *out, double a, long long *cyc)
{
    double v = threadIdx.x + 0.25;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = __dadd_rn((double)__double2int_rz(v), a);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// I2F.F64 -> DADD only... measured via int chain: i = f2i(i2f(i) + a)
__global__ void k_imad(int *out, int a, int b, long long *cyc)
{
    int v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = v * a + b;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// dependent 32-byte gathers: next index = low bits of the loaded value; all lanes the same address (broadcast)
// or lane-distinct sectors
template <int BYTES>
__global__ void k_chase(const double4 *tab, int mask, int lane_stride, double *out, long long *cyc, int n)
{
    int idx = (threadIdx.x * lane_stride) & mask;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double4 v;
        if (BYTES == 32)
            asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(tab + idx));
        else {
            double2 w;
            asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(w.x), "=d"(w.y) : "l"(tab + idx));
            v.x = w.x; v.y = w.y; v.z = 0; v.w = 0;
        }
        idx = (__double_as_longlong(v.x) + threadIdx.x * lane_stride) & mask;
        acc += v.y;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc + idx; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// the step chain itself (hoisted cell, constant-divisor quotients), one warp, table small enough for L1 or not
__global__ void k_step(const double4 *tab, int nxm1, int nym1, double st, double dx, double rdx, double *out, long long *cyc, int n)
{
    double x = 10.0 + 0.001 * threadIdx.x, y = 1.0, kx = 0.01, ky = 1.0 / st * 0.9;
    int ind = (int)y * nxm1 + (int)x;
    long long t0 = clock64();
    int i = 0;
    for (; i < n; ++i) {
        double4 h;
        asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(h.x), "=d"(h.y), "=d"(h.z), "=d"(h.w) : "l"(tab + ind));
        const double vx = __dadd_rn(h.x, __dmul_rn(kx, st)), vy = __dadd_rn(h.y, __dmul_rn(ky, st));
        const double qx = __dmul_rn(vx, rdx), qy = __dmul_rn(vy, rdx);
        const double rx = __fma_rn(-qx, dx, vx), ry = __fma_rn(-qy, dx, vy);
        x = __dadd_rn(x, __fma_rn(rx, rdx, qx)); y = __dadd_rn(y, __fma_rn(ry, rdx, qy));
        kx = __dadd_rn(kx, h.z); ky = __dadd_rn(ky, h.w);
        const int ix = __double2int_rz(x), iy = __double2int_rz(y);
        if ((unsigned)ix >= (unsigned)nxm1 || (unsigned)iy >= (unsigned)nym1) break;
        ind = iy * nxm1 + ix;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y + kx + ky; if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = i; }
}

// the same chain with the ray climbing along the table's FAST axis (the y-major copy): the next cells are the
// next 32-byte sectors of the same or the next line.  PF = 0: demand loads only; 1: prefetch.global.L1 of the
// cells +-2 ahead; 2: a real (discarded) 8-byte load of the cell +2 ahead; 3: prefetch.global.L2::evict_last
template <int PF>
__global__ void k_step_fast_axis(const double4 *tab, int nxm1, int nym1, double st, double dx, double rdx, double *out, long long *cyc, int n)
{
    double x = 10.0 + 0.001 * threadIdx.x, y = 1.0, kx = 0.01, ky = 1.0 / st * 0.9;
    int ind = (int)x * nym1 + (int)y;
    long long t0 = clock64();
    int i = 0;
    double sink = 0;
    for (; i < n; ++i) {
        double4 h;
        asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(h.x), "=d"(h.y), "=d"(h.z), "=d"(h.w) : "l"(tab + ind));
        if (PF == 1) {
            asm volatile("prefetch.global.L1 [%0+64];" :: "l"(tab + ind));
            asm volatile("prefetch.global.L1 [%0+-64];" :: "l"(tab + ind));
        } else if (PF == 2) {
            double t;
            asm volatile("ld.global.nc.f64 %0, [%1+64];" : "=d"(t) : "l"(tab + ind));
            sink = t;
        } else if (PF == 3) {
            asm volatile("prefetch.global.L1 [%0+128];" :: "l"(tab + ind));
        }
        const double vx = __dadd_rn(h.x, __dmul_rn(kx, st)), vy = __dadd_rn(h.y, __dmul_rn(ky, st));
        const double qx = __dmul_rn(vx, rdx), qy = __dmul_rn(vy, rdx);
        const double rx = __fma_rn(-qx, dx, vx), ry = __fma_rn(-qy, dx, vy);
        x = __dadd_rn(x, __fma_rn(rx, rdx, qx)); y = __dadd_rn(y, __fma_rn(ry, rdx, qy));
        kx = __dadd_rn(kx, h.z); ky = __dadd_rn(ky, h.w);
        const int ix = __double2int_rz(x), iy = __double2int_rz(y);
        if ((unsigned)ix >= (unsigned)nxm1 || (unsigned)iy >= (unsigned)nym1) break;
        ind = ix * nym1 + iy;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y + kx + ky + sink; if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = i; }
}

// Does an FP64 instruction hold the scheduler's dispatch port for two cycles?  Every warp runs NF independent DFMA
// chains and NS independent FFMA chains per iteration, 16 warps per scheduler.  If dispatch costs one cycle per
// instruction whatever its pipe, an iteration takes max(2 NF, NF + NS) cycles per warp and scheduler (the FP64 pipe
// is half rate); if an FP64 instruction blocks dispatch while it drains into its 16-lane pipe, 2 NF + NS.
template <int NF, int NS>
__global__ void __launch_bounds__(1024) k_mix(double *out, int iters, double a, double b, float fa, float fb)
{
    double d[NF > 0 ? NF : 1];
    float f[NS > 0 ? NS : 1];
#pragma unroll
    for (int k = 0; k < NF; ++k) d[k] = threadIdx.x + k;
#pragma unroll
    for (int k = 0; k < NS; ++k) f[k] = threadIdx.x + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < (NF > NS ? NF : NS); ++k) {
            if (k < NF) d[k] = __fma_rn(d[k], a, b);
            if (k < NS) f[k] = __fmaf_rn(f[k], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += d[k];
#pragma unroll
    for (int k = 0; k < NS; ++k) s += f[k];
    if (s == 123.456) out[threadIdx.x] = s;
}
template <int NF, int NS>
static void run_mix(double *out, int sms, double ghz)
{
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<NF, NS><<<sms * 2, 1024>>>(out, 100, 1.0000001, 0.5, 1.0000001f, 0.5f);
    cudaEventRecord(e0);
    k_mix<NF, NS><<<sms * 2, 1024>>>(out, iters, 1.0000001, 0.5, 1.0000001f, 0.5f);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9 / (16.0 * iters);     // 64 warps/SM = 16 per scheduler
    printf("%2d DFMA + %2d FFMA per iteration, 16 warps/scheduler: %6.1f cycles per warp-iteration "
           "(one dispatch cycle each: %d; FP64 holds the port two: %d)\n", NF, NS, cyc, (2 * NF > NF + NS ? 2 * NF : NF + NS), 2 * NF + NS);
}

int main()
{
    double *out; long long *cyc; int *iout;
    CK(cudaMalloc(&out, 4096)); CK(cudaMalloc(&iout, 4096)); CK(cudaMallocManaged(&cyc, 64));
    auto report = [&](const char *name, double per) { printf("%-56s %8.1f cycles\n", name, per); };
    k_dadd<<<1, 32>>>(out, 1.5, cyc); CK(cudaDeviceSynchronize()); report("DADD dependent chain, per op", (double)*cyc / N);
    k_dmul<<<1, 32>>>(out, 1.0000001, cyc); CK(cudaDeviceSynchronize()); report("DMUL dependent chain, per op", (double)*cyc / N);
    k_dfma<<<1, 32>>>(out, 1.0000001, 0.5, cyc); CK(cudaDeviceSynchronize()); report("DFMA dependent chain, per op", (double)*cyc / N);
    k_f2i<<<1, 32>>>(out, 0.25, cyc); CK(cudaDeviceSynchronize()); report("F2I.F64.TRUNC + I2F.F64 + DADD, per round", (double)*cyc / N);
    k_imad<<<1, 32>>>(iout, 3, 7, cyc); CK(cudaDeviceSynchronize()); report("IMAD dependent chain, per op", (double)*cyc / N);

    // gather tables: index stored in .x as an integer bit pattern (pseudo-random walk)
    for (int logn : {8, 12, 18, 21}) {             // 8 KB (L1), 128 KB (L1), 8 MB (L2), 64 MB (L2)
        const int n = 1 << logn;
        double4 *h = (double4 *)malloc((size_t)n * sizeof(double4));
        unsigned s = 12345;
        for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; long long nx = (s >> 8) & (n - 1); h[i].x = *(double *)&nx; h[i].y = 1.0; h[i].z = h[i].w = 0; }
        double4 *d; CK(cudaMalloc(&d, (size_t)n * sizeof(double4)));
        CK(cudaMemcpy(d, h, (size_t)n * sizeof(double4), cudaMemcpyHostToDevice));
        const int steps = 2000;
        for (int rep = 0; rep < 2; ++rep) {        // second pass: warm caches
            k_chase<32><<<1, 32>>>(d, n - 1, 0, out, cyc, steps); CK(cudaDeviceSynchronize());
        }
        char nm[128]; snprintf(nm, sizeof nm, "dependent 32 B gather, warp-uniform address, %d KB table", n / 32);
        report(nm, (double)*cyc / steps);
        for (int rep = 0; rep < 2; ++rep) { k_chase<32><<<1, 32>>>(d, n - 1, 1, out, cyc, steps); CK(cudaDeviceSynchronize()); }
        snprintf(nm, sizeof nm, "dependent 32 B gather, 32 distinct sectors, %d KB table", n / 32);
        report(nm, (double)*cyc / steps);
        for (int rep = 0; rep < 2; ++rep) { k_chase<16><<<1, 32>>>(d, n - 1, 0, out, cyc, steps); CK(cudaDeviceSynchronize()); }
        snprintf(nm, sizeof nm, "dependent 16 B gather, warp-uniform address, %d KB table", n / 32);
        report(nm, (double)*cyc / steps);
        cudaFree(d); free(h);
    }
    // the step chain on a uniform table: 1000 x 1000 cells of 32 B (32 MB, L2-resident after the first pass)
    {
        const int nxm1 = 1000, nym1 = 1000; const size_t n = (size_t)nxm1 * nym1;
        double4 *h = (double4 *)malloc(n * sizeof(double4));
        for (size_t i = 0; i < n; ++i) { h[i].x = 1e-9; h[i].y = -1e-9; h[i].z = 1e-9; h[i].w = -1e-7; }
        double4 *d; CK(cudaMalloc(&d, n * sizeof(double4)));
        CK(cudaMemcpy(d, h, n * sizeof(double4), cudaMemcpyHostToDevice));
        for (int rep = 0; rep < 3; ++rep) { k_step<<<1, 32>>>(d, nxm1, nym1, 0.9, 1.0, 1.0, out, cyc, 100000); CK(cudaDeviceSynchronize()); }
        printf("%-56s %8.1f cycles (%lld steps, a new row = a new sector every ~step)\n", "ray step chain, 1 warp, 32 MB table (rows)", (double)cyc[0] / cyc[1], cyc[1]);
        for (int rep = 0; rep < 3; ++rep) { k_step_fast_axis<0><<<1, 32>>>(d, nxm1, nym1, 0.9, 1.0, 1.0, out, cyc, 100000); CK(cudaDeviceSynchronize()); }
        printf("%-56s %8.1f cycles (%lld steps)\n", "ray step chain along the fast axis, no prefetch", (double)cyc[0] / cyc[1], cyc[1]);
        for (int rep = 0; rep < 3; ++rep) { k_step_fast_axis<1><<<1, 32>>>(d, nxm1, nym1, 0.9, 1.0, 1.0, out, cyc, 100000); CK(cudaDeviceSynchronize()); }
        printf("%-56s %8.1f cycles (%lld steps)\n", "  + prefetch.global.L1 of cells +-2", (double)cyc[0] / cyc[1], cyc[1]);
        for (int rep = 0; rep < 3; ++rep) { k_step_fast_axis<2><<<1, 32>>>(d, nxm1, nym1, 0.9, 1.0, 1.0, out, cyc, 100000); CK(cudaDeviceSynchronize()); }
        printf("%-56s %8.1f cycles (%lld steps)\n", "  + discarded 8-byte load of cell +2", (double)cyc[0] / cyc[1], cyc[1]);
        for (int rep = 0; rep < 3; ++rep) { k_step_fast_axis<3><<<1, 32>>>(d, nxm1, nym1, 0.9, 1.0, 1.0, out, cyc, 100000); CK(cudaDeviceSynchronize()); }
        printf("%-56s %8.1f cycles (%lld steps)\n", "  + prefetch.global.L1 of cell +4 (next line)", (double)cyc[0] / cyc[1], cyc[1]);
        cudaFree(d); free(h);
    }
    {
        int sms = 0, khz = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        const double ghz = khz * 1e-6;
        run_mix<8, 0>(out, sms, ghz);
        run_mix<0, 8>(out, sms, ghz);
        run_mix<8, 8>(out, sms, ghz);
        run_mix<4, 12>(out, sms, ghz);
        run_mix<8, 16>(out, sms, ghz);
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock (attribute): %d kHz\n", clk);
    return 0;
}
