// sidtrace_b200.cu -- C-ABI shared library (include/sidtrace_b200.h) over the
// sm_100a kernels in trace_kernels.cuh.  No Python / torch types anywhere.
//
// Build (see rayrtrace_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false \
//        -Xcompiler -fPIC -shared -o libsidtrace_b200.so sidtrace_b200.cu
#include "../../include/sidtrace_b200.h"
#include "trace_kernels.cuh"
#include "launch_kernels.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

constexpr int kCounters = 64;
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? STB_ERR_NOMEM : STB_ERR_CUDA,  \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return STB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return STB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {   // pinned
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return STB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        CU(cudaMallocHost(&p, want));
        cap = want;
        return STB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

// One hoisted table (trace_kernels.cuh, HOIST) per global step value the field has been traced with.
struct HoistEntry {
    uint64_t st_bits = 0;
    bool ymajor = false;                  // layout of the table = layout of the cells it was built from
    double4 *tab = nullptr, *tab_alloc = nullptr;
    cudaEvent_t ready = nullptr;          // recorded on the building stream after hoist_kernel
    uint64_t last_use = 0;
};
constexpr int kHoistSlots = 4;

struct stb_field_s {
    int device = 0;
    int nx = 0, ny = 0;
    double2 *cells = nullptr;             // interleaved {ax, ay}, (ny-1)*(nx-1), x-major (the reference's layout)
    double2 *cells_alloc = nullptr;       // its allocation (padded)
    double2 *cells_y = nullptr;           // the y-major copy, built on first use (STB_F_YMAJOR)
    double2 *cells_y_alloc = nullptr;
    cudaEvent_t cells_y_ready = nullptr;
    unsigned long long *counter = nullptr;   // kCounters queue heads, used round-robin; a slot is handed out
                                             // again only after the launch that used it last has finished
                                             // (counter_ev), so launches in flight never share one
    cudaEvent_t counter_ev[kCounters] = {};
    unsigned launch_seq = 0;              // guarded by mu
    HoistEntry hoist[kHoistSlots];
    uint64_t hoist_clock = 0;
    size_t hoist_max_bytes = 64u << 20;   // larger tables would fall out of L2: keep the 16-byte cells
    int num_sms = 0;
    std::mutex mu;                        // guards the workspaces below
    DevBuf ws_in, ws_out, ws_traj, ws_hist;
    HostBuf pin;
};

// shared with rk_b200.cu (same library, not part of the public headers)
extern "C" int stb_internal_fail(int code, const char *msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
extern "C" void stb_internal_count_launch(void) { g_launches++; }

extern "C" {

int stb_abi_version(void) { return STB_ABI_VERSION; }

const char *stb_last_error(void) { return g_err; }

int stb_device_count(int *count)
{
    if (!count) return fail(STB_ERR_ARG, "count is NULL");
    CU(cudaGetDeviceCount(count));
    return STB_OK;
}

int stb_set_device(int device)
{
    CU(cudaSetDevice(device));
    return STB_OK;
}

int64_t stb_launch_count(void) { return g_launches.load(); }

int stb_field_create(const double *ax, const double *ay, int nx, int ny, int on_device, stb_field_t *out)
{
    if (!ax || !ay || !out) return fail(STB_ERR_ARG, "stb_field_create: NULL argument");
    if (nx < 2 || ny < 2) return fail(STB_ERR_ARG, "stb_field_create: nx, ny must be >= 2 (got %d, %d)", nx, ny);
    const long long ncell = (long long)(nx - 1) * (long long)(ny - 1);
    if (ncell > 0x7fffffffLL) return fail(STB_ERR_RANGE, "stb_field_create: (nx-1)*(ny-1) = %lld exceeds int32", ncell);

    stb_field_s *f = new (std::nothrow) stb_field_s;
    if (!f) return fail(STB_ERR_NOMEM, "stb_field_create: host allocation failed");
    f->nx = nx; f->ny = ny;
    cudaError_t e = cudaGetDevice(&f->device);
    if (e != cudaSuccess) { delete f; return fail(STB_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&f->num_sms, cudaDevAttrMultiProcessorCount, f->device);
    if (const char *e = getenv("STB_HOIST_MAX_MB")) f->hoist_max_bytes = (size_t)atoll(e) << 20;

    const size_t bytes = (size_t)ncell * sizeof(double);
    double *tmp = nullptr;   // staging for host input: [ax | ay]
    const double *dax = ax, *day = ay;
    int rc = STB_OK;
    do {
        // every cell table carries kTablePadCells cells of padding on either side (prefetch_neighbours)
        if ((e = cudaMalloc(&f->cells_alloc, 2 * bytes + 2 * stb::kTablePadCells * sizeof(double2))) != cudaSuccess) break;
        f->cells = f->cells_alloc + stb::kTablePadCells;
        if ((e = cudaMalloc(&f->counter, kCounters * sizeof(unsigned long long))) != cudaSuccess) break;
        if (!on_device) {
            if ((e = cudaMalloc(&tmp, 2 * bytes)) != cudaSuccess) break;
            if ((e = cudaMemcpy(tmp, ax, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
            if ((e = cudaMemcpy(tmp + ncell, ay, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
            dax = tmp; day = tmp + ncell;
        }
        const int threads = 256;
        long long blocks = (ncell + threads - 1) / threads;
        if (blocks > 148 * 16) blocks = 148 * 16;
        stb::interleave_kernel<<<(int)blocks, threads>>>(dax, day, f->cells, ncell);
        g_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
    } while (0);
    if (tmp) cudaFree(tmp);
    if (e != cudaSuccess) {
        rc = fail(e == cudaErrorMemoryAllocation ? STB_ERR_NOMEM : STB_ERR_CUDA,
                  "stb_field_create: %s", cudaGetErrorString(e));
        if (f->cells_alloc) cudaFree(f->cells_alloc);
        if (f->counter) cudaFree(f->counter);
        delete f;
        return rc;
    }
    *out = f;
    return STB_OK;
}

int stb_field_destroy(stb_field_t f)
{
    if (!f) return STB_OK;
    cudaFree(f->cells_alloc);
    if (f->cells_y_alloc) cudaFree(f->cells_y_alloc);
    if (f->cells_y_ready) cudaEventDestroy(f->cells_y_ready);
    cudaFree(f->counter);
    for (auto &h : f->hoist) {
        if (h.tab_alloc) cudaFree(h.tab_alloc);
        if (h.ready) cudaEventDestroy(h.ready);
    }
    for (auto &e : f->counter_ev) if (e) cudaEventDestroy(e);
    f->ws_in.release(); f->ws_out.release(); f->ws_traj.release(); f->ws_hist.release();
    f->pin.release();
    delete f;
    return STB_OK;
}

int stb_field_shape(stb_field_t f, int *nx, int *ny)
{
    if (!f) return fail(STB_ERR_ARG, "stb_field_shape: NULL field");
    if (nx) *nx = f->nx;
    if (ny) *ny = f->ny;
    return STB_OK;
}

int stb_field_peek(stb_field_t f, int i, int j, int which, double *out)
{
    if (!f || !out) return fail(STB_ERR_ARG, "stb_field_peek: NULL argument");
    if (i < 0 || i >= f->ny - 1 || j < 0 || j >= f->nx - 1 || (which != 0 && which != 1))
        return fail(STB_ERR_ARG, "stb_field_peek: index (%d, %d, %d) outside (%d, %d)", i, j, which, f->ny - 1, f->nx - 1);
    double2 c;
    CU(cudaMemcpy(&c, f->cells + (size_t)i * (f->nx - 1) + j, sizeof(c), cudaMemcpyDeviceToHost));
    *out = which ? c.y : c.x;
    return STB_OK;
}

int stb_check(int nx, int ny, int i, int j, const double *a, double *out)
{
    (void)ny;
    if (!a || !out) return fail(STB_ERR_ARG, "stb_check: NULL argument");
    // sidtracemodule.c:66 -- a[i*nx + j]; a plain host-memory probe, no compute.
    *out = a[(long long)i * nx + j];
    return STB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
namespace {

using stb::KParams;

// STB_DUAL=0 in the environment keeps the one-ray-per-thread endpoint kernel on the static grid (A/B measurements)
const bool g_dual = [] { const char *e = getenv("STB_DUAL"); return !(e && e[0] == '0'); }();

struct LaunchCtx {
    stb_field_s *f;
    bool refill;
    cudaStream_t s;
};

// The actual launch of one kernel instantiation.
template <bool TRAJ, bool EXTRAS, bool HIST, bool DIVC, bool TILED, bool CROSS, bool LOOKUP, bool STRAY, bool HOIST>
int launch_one(const KParams &p0, const LaunchCtx &c)
{
    cudaStream_t s = c.s;
    // static tiled trajectories over 256-ray tiles: 256-thread CTAs, one 8 KB row per CTA-step
    if constexpr (TRAJ && TILED && !EXTRAS && !HIST && !CROSS && !STRAY) {
        if (!c.refill && p0.tile == 256 && p0.nrays > 0) {
            const long long blocks = ((p0.perm ? p0.nslots : p0.nrays) + 255) / 256;
            if (blocks > 0x7fffffffLL) return fail(STB_ERR_ARG, "too many rays for a static grid");
            stb::trace_kernel<true, false, false, DIVC, false, true, false, false, false, 256, HOIST><<<(unsigned)blocks, 256, 0, s>>>(p0);
            g_launches++;
            CU(cudaGetLastError());
            return STB_OK;
        }
    }
    const size_t smem = HIST ? (size_t)p0.nbins * sizeof(double) * (p0.hist_table ? 3 : 1) : 0;
    if (p0.nrays <= 0) return STB_OK;
    const long long nwork = p0.perm ? p0.nslots : p0.nrays;     // slots to hand out
    // step counts / end states only, static grid: two rays per thread (endpoint_dual_kernel)
    if constexpr (!TRAJ && !EXTRAS && !HIST && !TILED && !CROSS && !STRAY) {
        if (!c.refill && g_dual) {
            const long long blocks = (nwork + 2 * stb::kBlock - 1) / (2 * stb::kBlock);
            if (blocks > 0x7fffffffLL) return fail(STB_ERR_ARG, "too many rays for a static grid");
            stb::endpoint_dual_kernel<DIVC, HOIST><<<(unsigned)blocks, stb::kBlock, 0, s>>>(p0);
            g_launches++;
            CU(cudaGetLastError());
            return STB_OK;
        }
    }
    long long blocks_all = (nwork + stb::kBlock - 1) / stb::kBlock;
    if (c.refill) {
        auto kern = stb::trace_kernel<TRAJ, EXTRAS, HIST, DIVC, true, TILED, CROSS, STRAY, LOOKUP, stb::kBlock, HOIST>;
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, stb::kBlock, smem));
        if (per_sm < 1) per_sm = 1;
        long long blocks = (long long)c.f->num_sms * per_sm;
        if (blocks > blocks_all) blocks = blocks_all;
        // queue head: one of kCounters slots; a slot is reused only after the launch that used it last
        // has finished, so launches in flight (other streams, other host threads) never share one
        stb_field_s *f = c.f;
        const int k = (int)(f->launch_seq++ % kCounters);
        if (f->counter_ev[k]) CU(cudaEventSynchronize(f->counter_ev[k]));
        else CU(cudaEventCreateWithFlags(&f->counter_ev[k], cudaEventDisableTiming));
        KParams p = p0;
        p.counter = f->counter + k;
        CU(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), s));
        kern<<<(unsigned)blocks, stb::kBlock, smem, s>>>(p);
        CU(cudaEventRecord(f->counter_ev[k], s));
    } else {
        auto kern = stb::trace_kernel<TRAJ, EXTRAS, HIST, DIVC, false, TILED, CROSS, STRAY, LOOKUP, stb::kBlock, HOIST>;
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (blocks_all > 0x7fffffffLL) return fail(STB_ERR_ARG, "too many rays for a static grid");
        kern<<<(unsigned)blocks_all, stb::kBlock, smem, s>>>(p0);
    }
    g_launches++;
    CU(cudaGetLastError());
    return STB_OK;
}

// Step flavour: per-ray steps (STRAY); one global step with the hoisted per-cell products (HOIST);
// or one global step read from the constant bank (lean variants only, when hoisting is off).
template <bool TRAJ, bool EXTRAS, bool HIST, bool DIVC, bool TILED = false, bool CROSS = false, bool LOOKUP = false>
int launch_variant(const KParams &p, const LaunchCtx &c)
{
    if (!p.st_ray) {
        if (p.field4) return launch_one<TRAJ, EXTRAS, HIST, DIVC, TILED, CROSS, LOOKUP, false, true>(p, c);
        if constexpr (!EXTRAS && !HIST && !CROSS)
            return launch_one<TRAJ, EXTRAS, HIST, DIVC, TILED, CROSS, LOOKUP, false, false>(p, c);
    }
    return launch_one<TRAJ, EXTRAS, HIST, DIVC, TILED, CROSS, LOOKUP, true, false>(p, c);
}

template <bool TRAJ, bool EXTRAS, bool HIST>
int launch_div(const KParams &p, bool divc, const LaunchCtx &c)
{
    return divc ? launch_variant<TRAJ, EXTRAS, HIST, true>(p, c)
                : launch_variant<TRAJ, EXTRAS, HIST, false>(p, c);
}

template <bool TRAJ, bool TILED>
int launch_cross(const KParams &p, bool divc, const LaunchCtx &c)
{
    return divc ? launch_variant<TRAJ, true, false, true, TILED, true>(p, c)
                : launch_variant<TRAJ, true, false, false, TILED, true>(p, c);
}

// Table of the cell-only products for the launch's global step: built once per (field, st) on the
// launching stream, kept in a small LRU cache on the field.  Caller holds f->mu.
// The y-major copy of the cells, built once on the launching stream.  Caller holds f->mu.
int ensure_cells_y(stb_field_s *f, cudaStream_t s, const double2 **out)
{
    if (!f->cells_y) {
        const long long ncell = (long long)(f->nx - 1) * (long long)(f->ny - 1);
        CU(cudaMalloc(&f->cells_y_alloc, ((size_t)ncell + 2 * stb::kTablePadCells) * sizeof(double2)));
        f->cells_y = f->cells_y_alloc + stb::kTablePadCells;
        CU(cudaEventCreateWithFlags(&f->cells_y_ready, cudaEventDisableTiming));
        const dim3 grid((f->nx - 1 + 31) / 32, (f->ny - 1 + 31) / 32), block(32, 8);
        stb::transpose_cells_kernel<double2><<<grid, block, 0, s>>>(f->cells, f->cells_y, f->nx - 1, f->ny - 1);
        g_launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(f->cells_y_ready, s));
    } else {
        CU(cudaStreamWaitEvent(s, f->cells_y_ready, 0));
    }
    *out = f->cells_y;
    return STB_OK;
}

int ensure_hoist(stb_field_s *f, double st, bool ymajor, const double2 *cells, cudaStream_t s, const double4 **out)
{
    uint64_t bits;
    memcpy(&bits, &st, sizeof(bits));
    HoistEntry *e = nullptr, *victim = nullptr;
    for (auto &h : f->hoist) {
        if (h.tab && h.st_bits == bits && h.ymajor == ymajor) { e = &h; break; }
        if (!victim || (victim->tab && (!h.tab || h.last_use < victim->last_use))) victim = &h;
    }
    if (!e) {
        e = victim;
        const long long ncell = (long long)(f->nx - 1) * (long long)(f->ny - 1);
        if (e->tab) CU(cudaDeviceSynchronize());        // kernels that still read the evicted table
        else {
            CU(cudaMalloc(&e->tab_alloc, ((size_t)ncell + 2 * stb::kTablePadCells) * sizeof(double4)));
            e->tab = e->tab_alloc + stb::kTablePadCells;
        }
        if (!e->ready) CU(cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming));
        long long blocks = (ncell + 255) / 256;
        if (blocks > (long long)f->num_sms * 16) blocks = (long long)f->num_sms * 16;
        stb::hoist_kernel<<<(unsigned)blocks, 256, 0, s>>>(cells, st, st * st, e->tab, ncell);
        g_launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->ready, s));
        e->st_bits = bits;
        e->ymajor = ymajor;
    } else {
        CU(cudaStreamWaitEvent(s, e->ready, 0));         // built on another stream, perhaps
    }
    e->last_use = ++f->hoist_clock;
    *out = e->tab;
    return STB_OK;
}

int launch_trace(stb_field_s *f, KParams &p, bool divc, bool refill, bool tiled, bool hoist, bool ymajor, cudaStream_t s)
{
    const LaunchCtx c{f, refill, s};
    p.field4 = nullptr;
    if (ymajor && p.nrays > 0) {
        int rc = ensure_cells_y(f, s, &p.field);
        if (rc != STB_OK) return rc;
        p.csx = p.nym1; p.csy = 1;
    }
    // plain trajectory launches are bound by HBM stores, not by FP64 issue: the 32-byte hoisted cells only
    // add L2 traffic there (6770 vs 6810 GB/s, profiles/r02/l2_policy_sweep.txt) -- they keep the 16-byte cells
    const bool lean_traj = p.traj && !p.hist && !p.cross_axis && !(p.turn || p.turn_state || p.lax_end || p.ph_end);
    if (hoist && !lean_traj && !p.st_ray && p.nrays > 0) {
        const size_t bytes = (size_t)(f->nx - 1) * (size_t)(f->ny - 1) * sizeof(double4);
        if (bytes <= f->hoist_max_bytes) {
            int rc = ensure_hoist(f, p.st, ymajor && p.nrays > 0, p.field, s, &p.field4);
            if (rc != STB_OK) return rc;
        }
    }
    if (p.cross_axis) {
        if (p.hist) return fail(STB_ERR_ARG, "the crossing output cannot be combined with the histogram output");
        if (p.lookup_ph) {
            if (p.traj || tiled) return fail(STB_ERR_ARG, "the equal-phase lookup pass takes no trajectory output and no tiled schedule");
            return divc ? launch_variant<false, true, false, true, false, true, true>(p, c)
                        : launch_variant<false, true, false, false, false, true, true>(p, c);
        }
        if (p.traj) return tiled ? launch_cross<true, true>(p, divc, c) : launch_cross<true, false>(p, divc, c);
        return tiled ? launch_cross<false, true>(p, divc, c) : launch_cross<false, false>(p, divc, c);
    }
    if (tiled) {
        // warp-tile schedule: plain trajectories and/or fused running sums; the histogram keeps
        // the ray-granular kernels
        const bool extras = p.turn || p.turn_state || p.lax_end || p.ph_end;
        if (p.hist) return fail(STB_ERR_ARG, "STB_F_TRAJ_TILED cannot be combined with the histogram output");
        const int key = (p.traj ? 4 : 0) | (extras ? 2 : 0) | (divc ? 1 : 0);
        switch (key) {
            case 0: return launch_variant<false, false, false, false, true>(p, c);
            case 1: return launch_variant<false, false, false, true, true>(p, c);
            case 2: return launch_variant<false, true, false, false, true>(p, c);
            case 3: return launch_variant<false, true, false, true, true>(p, c);
            case 4: return launch_variant<true, false, false, false, true>(p, c);
            case 5: return launch_variant<true, false, false, true, true>(p, c);
            case 6: return launch_variant<true, true, false, false, true>(p, c);
            default: return launch_variant<true, true, false, true, true>(p, c);
        }
    }
    const bool traj = p.traj != nullptr;
    const bool extras = p.turn || p.turn_state || p.lax_end || p.ph_end;
    const bool hist = p.hist != nullptr;
    const int key = (traj ? 4 : 0) | (extras ? 2 : 0) | (hist ? 1 : 0);
    switch (key) {
        case 0: return launch_div<false, false, false>(p, divc, c);
        case 1: return launch_div<false, false, true>(p, divc, c);
        case 2: return launch_div<false, true, false>(p, divc, c);
        case 3: return launch_div<false, true, true>(p, divc, c);
        case 4: return launch_div<true, false, false>(p, divc, c);
        case 5: return launch_div<true, false, true>(p, divc, c);
        case 6: return launch_div<true, true, false>(p, divc, c);
        default: return launch_div<true, true, true>(p, divc, c);
    }
}

// Device alias of a pinned (page-locked, mapped) host pointer, or nullptr when the memory is
// pageable / unknown to CUDA.  With unified addressing the alias equals the host pointer.
void *mapped_alias(const void *host)
{
    if (!host) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost) return nullptr;
    return a.devicePointer;
}

// the constant-divisor sequence is exact only while nothing under/overflows
bool divisor_in_guarded_range(double d)
{
    const double a = std::fabs(d);
    return std::isfinite(d) && a >= std::ldexp(1.0, -400) && a <= std::ldexp(1.0, 400);
}

int fill_params(stb_field_t f, const stb_batch_t *b, KParams &p, bool &divc)
{
    if (b->L < 1 || b->L > 0x7fffffffLL) return fail(STB_ERR_ARG, "stb_trace_batch: L = %lld out of range", (long long)b->L);
    if (b->nrays < 0) return fail(STB_ERR_ARG, "stb_trace_batch: nrays < 0");
    if (b->nrays > 0 && (!b->x0 || !b->y0 || !b->kx0 || !b->ky0))
        return fail(STB_ERR_ARG, "stb_trace_batch: NULL launch-state pointer");
    long long stride = b->traj_ray_stride ? b->traj_ray_stride : 4 * b->L;
    if (b->traj && (stride < 4 * b->L || (stride & 3)))
        return fail(STB_ERR_ARG, "stb_trace_batch: traj_ray_stride %lld must be a multiple of 4 and >= 4*L", stride);
    if (b->traj && ((uintptr_t)b->traj & 31) && !(b->flags & STB_F_TRAJ_HOST))
        return fail(STB_ERR_ARG, "stb_trace_batch: traj must be 32-byte aligned");
    if (b->hist && (b->nbins < 1 || b->nbins > 16384))
        return fail(STB_ERR_ARG, "stb_trace_batch: nbins = %d outside [1, 16384]", b->nbins);
    if (b->hist && !(b->r_lcs > 0 && b->rho_max > 0))
        return fail(STB_ERR_ARG, "stb_trace_batch: r_lcs and rho_max must be > 0 for the histogram");
    p.field = f->cells;
    p.nxm1 = f->nx - 1; p.nym1 = f->ny - 1; p.ncell = (f->nx - 1) * (f->ny - 1);
    p.csx = 1; p.csy = f->nx - 1;
    p.fx = (double)(f->nx - 1); p.fy = (double)(f->ny - 1);
    p.nrays = b->nrays; p.L = (int)b->L; p.Lm1 = p.L - 1;
    p.x0 = b->x0; p.y0 = b->y0; p.kx0 = b->kx0; p.ky0 = b->ky0;
    p.st = b->st; p.sst = b->st * b->st; p.st_ray = b->st_ray;
    p.dx = b->dx; p.dy = b->dy; p.rdx = 1.0 / b->dx; p.rdy = 1.0 / b->dy;
    p.traj = b->traj; p.stride = stride;
    p.tile = b->traj_tile ? b->traj_tile : 256;
    if (p.tile != 32 && p.tile != 64 && p.tile != 128 && p.tile != 256)
        return fail(STB_ERR_ARG, "stb_trace_batch: traj_tile must be 32, 64, 128 or 256 (got %d)", b->traj_tile);
    p.perm = nullptr; p.tile_off = nullptr; p.nslots = 0;
    if (b->traj_tile_off) {
        // packed tiled trajectories (stb_traj_plan)
        if (!(b->flags & STB_F_TRAJ_TILED) || !b->traj_perm)
            return fail(STB_ERR_ARG, "stb_trace_batch: a packed plan needs STB_F_TRAJ_TILED, traj_perm and traj_tile_off");
        const long long want = (b->nrays + p.tile - 1) / p.tile * p.tile;
        if (b->traj_nslots != want)
            return fail(STB_ERR_ARG, "stb_trace_batch: traj_nslots = %lld, expected ceil(nrays/T)*T = %lld", (long long)b->traj_nslots, want);
        if (b->flags & STB_F_HOST_IO)
            return fail(STB_ERR_ARG, "stb_trace_batch: a packed plan lives in device memory (no STB_F_HOST_IO)");
        p.perm = b->traj_perm; p.tile_off = (const long long *)b->traj_tile_off; p.nslots = b->traj_nslots;
    } else if (b->traj_perm) {
        // tracing order only: slot -> ray; tiled trajectories are laid out by slot
        const long long want = (b->flags & STB_F_TRAJ_TILED) ? (b->nrays + p.tile - 1) / p.tile * p.tile : b->nrays;
        if (b->traj_nslots < b->nrays || ((b->flags & STB_F_TRAJ_TILED) && b->traj && b->traj_nslots != want))
            return fail(STB_ERR_ARG, "stb_trace_batch: traj_nslots = %lld does not fit %lld rays (tiled trajectories: ceil(nrays/T)*T = %lld)",
                        (long long)b->traj_nslots, (long long)b->nrays, want);
        if (b->flags & (STB_F_HOST_IO | STB_F_TRAJ_HOST))
            return fail(STB_ERR_ARG, "stb_trace_batch: a tracing order (traj_perm) needs device I/O");
        p.perm = b->traj_perm; p.nslots = b->traj_nslots;
    }
    p.nsteps = b->nsteps; p.endst = b->end_state;
    p.turn = b->turn_index; p.turn_state = b->turn_state; p.lax_end = b->lax_end; p.ph_end = b->ph_end;
    p.hist = b->hist; p.nbins = b->nbins; p.amp = b->amp;
    p.xorg = b->xorg; p.yorg = b->yorg; p.r_lcs = b->r_lcs; p.rho_max = b->rho_max;
    p.r_rlcs = 1.0 / b->r_lcs; p.r_rhomax = 1.0 / b->rho_max;
    p.hist_dr = b->nbins > 0 ? b->rho_max * b->r_lcs / (double)b->nbins : 0.0;
    // a ray advances about one smallest cell per step (SURVEY 3.3 i): the annulus test (two compares) saves
    // the square root and both divisions of the flux label whenever the ray is still inside its bin; it
    // pays off as soon as a bin is about one step wide (failing it costs 2 of ~60 instructions)
    p.hist_annulus = b->hist && p.hist_dr >= 1.0 * std::fmin(std::fabs(b->dx), std::fabs(b->dy));
    p.hist_table = p.hist_annulus && b->nbins <= 2048;        // 3 x 16 KB of shared memory per CTA at most
    p.counter = nullptr;                 // assigned per persistent launch (launch_one)
    p.cross_axis = b->cross_axis; p.cross_val = b->cross_val;
    p.x0p = b->x0_phys; p.y0p = b->y0_phys;
    p.cross_count = b->cross_count; p.cross_index = b->cross_index; p.cross_flags = b->cross_flags;
    p.cross_state = b->cross_state;
    p.lookup_ph = b->lookup_ph; p.lookup_xy = b->lookup_xy;
    if (b->lookup_ph && (!b->cross_axis || !b->lookup_xy))
        return fail(STB_ERR_ARG, "stb_trace_batch: lookup_ph needs cross_axis != 0 and lookup_xy");
    if (b->cross_axis) {
        if (b->cross_axis != 1 && b->cross_axis != 2) return fail(STB_ERR_ARG, "stb_trace_batch: cross_axis must be 0, 1 (y) or 2 (x)");
        if (b->nrays > 0 && (!b->x0_phys || !b->y0_phys))
            return fail(STB_ERR_ARG, "stb_trace_batch: the crossing output needs x0_phys and y0_phys");
    }
    divc = !(b->flags & STB_F_DIV_IEEE) && divisor_in_guarded_range(b->dx) && divisor_in_guarded_range(b->dy);
    if (b->hist) divc = divc && divisor_in_guarded_range(b->r_lcs) && divisor_in_guarded_range(b->rho_max);
    return STB_OK;
}

}  // namespace

extern "C" {

int stb_trace_batch(stb_field_t f, const stb_batch_t *b, void *stream)
{
    if (!f || !b) return fail(STB_ERR_ARG, "stb_trace_batch: NULL argument");
    if (b->struct_size != sizeof(stb_batch_t))
        return fail(STB_ERR_ARG, "stb_trace_batch: struct_size %zu != %zu (ABI mismatch)", b->struct_size, sizeof(stb_batch_t));
    cudaStream_t s = (cudaStream_t)stream;
    const bool host_io = b->flags & STB_F_HOST_IO;
    const bool traj_host = b->flags & STB_F_TRAJ_HOST;
    const bool tiled = b->flags & STB_F_TRAJ_TILED;
    if ((b->flags & STB_F_REFILL) && (b->flags & STB_F_STATIC)) return fail(STB_ERR_ARG, "STB_F_REFILL and STB_F_STATIC are mutually exclusive");
    if (tiled && traj_host) return fail(STB_ERR_ARG, "STB_F_TRAJ_TILED and STB_F_TRAJ_HOST are mutually exclusive");
    const bool refill = (b->flags & STB_F_REFILL) ? true : (b->flags & STB_F_STATIC) ? false : !tiled;
    int dev = -1;
    CU(cudaGetDevice(&dev));
    if (dev != f->device)
        return fail(STB_ERR_ARG, "stb_trace_batch: the field lives on device %d but the current device is %d", f->device, dev);

    // one lock per field: the queue heads and the staging workspaces belong to it
    std::lock_guard<std::mutex> lk(f->mu);
    KParams p{};
    bool divc = false;
    int rc = fill_params(f, b, p, divc);
    if (rc != STB_OK) return rc;

    const bool hoist = !(b->flags & STB_F_NO_HOIST);
    const bool ymajor = b->flags & STB_F_YMAJOR;
    if (!host_io && !traj_host) return launch_trace(f, p, divc, refill, tiled, hoist, ymajor, s);

    const size_t n = (size_t)b->nrays;
    const size_t nb = n * sizeof(double);

    // ---- zero-copy host I/O -------------------------------------------------------------
    // With the warp-tile schedule every launch-state read and every per-ray result write is a
    // coalesced 128-256 B row, so when the caller's buffers are pinned the kernel reads and
    // writes them in place over PCIe, overlapped with the tracing itself: no staging copies,
    // no serial H2D -> kernel -> D2H chain.  Pageable buffers take the staged path below.
    if (host_io && tiled && !traj_host && !b->hist && n) {
        const void *hp[18] = {b->x0, b->y0, b->kx0, b->ky0, b->st_ray, b->amp,
                              b->nsteps, b->end_state, b->turn_index, b->lax_end, b->ph_end,
                              b->x0_phys, b->y0_phys, b->cross_count, b->cross_index, b->cross_flags, b->cross_state,
                              b->turn_state};
        void *dp[18];
        bool all = true;
        for (int k = 0; k < 18; ++k) {
            dp[k] = mapped_alias(hp[k]);
            if (hp[k] && !dp[k]) all = false;
        }
        if (all) {
            p.x0 = (const double *)dp[0]; p.y0 = (const double *)dp[1];
            p.kx0 = (const double *)dp[2]; p.ky0 = (const double *)dp[3];
            p.st_ray = (const double *)dp[4]; p.amp = (const double *)dp[5];
            p.nsteps = (int *)dp[6]; p.endst = (double *)dp[7];
            p.turn = (int *)dp[8]; p.lax_end = (double *)dp[9]; p.ph_end = (double *)dp[10];
            p.x0p = (const double *)dp[11]; p.y0p = (const double *)dp[12];
            p.cross_count = (int *)dp[13]; p.cross_index = (int *)dp[14]; p.cross_flags = (int *)dp[15];
            p.cross_state = (double *)dp[16];
            p.turn_state = (double *)dp[17];
            rc = launch_trace(f, p, divc, refill, tiled, hoist, ymajor, s);
            if (rc != STB_OK) return rc;
            CU(cudaStreamSynchronize(s));
            return STB_OK;
        }
    }

    if (host_io) {
        // ---- inputs: [x0 | y0 | kx0 | ky0 | st_ray? | amp?] ----
        const int nin = 4 + (b->st_ray ? 1 : 0) + (b->amp ? 1 : 0) + (b->x0_phys ? 1 : 0) + (b->y0_phys ? 1 : 0)
                        + (b->lookup_ph ? 2 : 0);
        rc = f->ws_in.ensure(nb * nin + 64);
        if (rc != STB_OK) return rc;
        double *din = (double *)f->ws_in.p;
        const double *src[8] = {b->x0, b->y0, b->kx0, b->ky0, b->st_ray, b->amp, b->x0_phys, b->y0_phys};
        const double **dst[8] = {&p.x0, &p.y0, &p.kx0, &p.ky0, &p.st_ray, &p.amp, &p.x0p, &p.y0p};
        int slot = 0;
        for (int k = 0; k < 8; ++k) {
            if (!src[k]) continue;
            double *d = din + (size_t)slot * n;
            if (n) CU(cudaMemcpyAsync(d, src[k], nb, cudaMemcpyHostToDevice, s));
            *dst[k] = d;
            ++slot;
        }
        if (b->lookup_ph) {
            double *d = din + (size_t)slot * n;
            if (n) CU(cudaMemcpyAsync(d, b->lookup_ph, 2 * nb, cudaMemcpyHostToDevice, s));
            p.lookup_ph = d;
        }
        // ---- outputs: [end 4n | lax n | ph n | cross_state 6n | nsteps, turn, cross_count/index/flags n(i32) each] ----
        rc = f->ws_out.ensure(nb * 20 + 5 * n * sizeof(int32_t) + 64);
        if (rc != STB_OK) return rc;
        double *dout = (double *)f->ws_out.p;
        p.endst = b->end_state ? dout : nullptr;
        p.lax_end = b->lax_end ? dout + 4 * n : nullptr;
        p.ph_end = b->ph_end ? dout + 5 * n : nullptr;
        p.cross_state = b->cross_state ? dout + 6 * n : nullptr;
        p.lookup_xy = b->lookup_xy ? dout + 12 * n : nullptr;
        p.turn_state = b->turn_state ? dout + 16 * n : nullptr;
        int32_t *iout = (int32_t *)(dout + 20 * n);
        p.nsteps = b->nsteps ? iout : nullptr;
        p.turn = b->turn_index ? iout + n : nullptr;
        p.cross_count = b->cross_count ? iout + 2 * n : nullptr;
        p.cross_index = b->cross_index ? iout + 3 * n : nullptr;
        p.cross_flags = b->cross_flags ? iout + 4 * n : nullptr;
        if (b->hist) {
            rc = f->ws_hist.ensure((size_t)b->nbins * sizeof(double));
            if (rc != STB_OK) return rc;
            CU(cudaMemcpyAsync(f->ws_hist.p, b->hist, (size_t)b->nbins * sizeof(double), cudaMemcpyHostToDevice, s));
            p.hist = (double *)f->ws_hist.p;
        }
    }
    if (traj_host && b->traj) {
        rc = f->ws_traj.ensure((size_t)b->nrays * (size_t)p.stride * sizeof(double));
        if (rc != STB_OK) return rc;
        // entries beyond state i stay untouched for the caller: start from the caller's bytes
        CU(cudaMemcpyAsync(f->ws_traj.p, b->traj, (size_t)b->nrays * (size_t)p.stride * sizeof(double), cudaMemcpyHostToDevice, s));
        p.traj = (double *)f->ws_traj.p;
    }

    rc = launch_trace(f, p, divc, refill, tiled, hoist, ymajor, s);
    if (rc != STB_OK) return rc;

    if (host_io && n) {
        if (b->end_state) CU(cudaMemcpyAsync(b->end_state, p.endst, 4 * nb, cudaMemcpyDeviceToHost, s));
        if (b->lax_end) CU(cudaMemcpyAsync(b->lax_end, p.lax_end, nb, cudaMemcpyDeviceToHost, s));
        if (b->ph_end) CU(cudaMemcpyAsync(b->ph_end, p.ph_end, nb, cudaMemcpyDeviceToHost, s));
        if (b->nsteps) CU(cudaMemcpyAsync(b->nsteps, p.nsteps, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (b->turn_index) CU(cudaMemcpyAsync(b->turn_index, p.turn, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (b->cross_state) CU(cudaMemcpyAsync(b->cross_state, p.cross_state, 6 * nb, cudaMemcpyDeviceToHost, s));
        if (b->lookup_xy) CU(cudaMemcpyAsync(b->lookup_xy, p.lookup_xy, 4 * nb, cudaMemcpyDeviceToHost, s));
        if (b->turn_state) CU(cudaMemcpyAsync(b->turn_state, p.turn_state, 4 * nb, cudaMemcpyDeviceToHost, s));
        if (b->cross_count) CU(cudaMemcpyAsync(b->cross_count, p.cross_count, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (b->cross_index) CU(cudaMemcpyAsync(b->cross_index, p.cross_index, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (b->cross_flags) CU(cudaMemcpyAsync(b->cross_flags, p.cross_flags, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    if (host_io && b->hist)
        CU(cudaMemcpyAsync(b->hist, p.hist, (size_t)b->nbins * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (traj_host && b->traj)
        CU(cudaMemcpyAsync(b->traj, p.traj, (size_t)b->nrays * (size_t)p.stride * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return STB_OK;
}

int stb_trace(stb_field_t f, double *x, double *y, double *kx, double *ky, int64_t L,
              double st, double dx, double dy, int64_t *nsteps)
{
    if (!f || !x || !y || !kx || !ky || !nsteps) return fail(STB_ERR_ARG, "stb_trace: NULL argument");
    if (L < 1 || L > 0x7fffffffLL) return fail(STB_ERR_ARG, "stb_trace: L = %lld out of range", (long long)L);
    int dev = -1;
    CU(cudaGetDevice(&dev));
    if (dev != f->device)
        return fail(STB_ERR_ARG, "stb_trace: the field lives on device %d but the current device is %d", f->device, dev);
    std::lock_guard<std::mutex> lk(f->mu);
    // One pinned, device-mapped workspace: [launch state 4 doubles | nsteps | pad] [trajectory L x 4].
    // The single-ray kernel reads the launch state and writes every state straight into it over
    // PCIe (posted stores, off the critical path), so a call is one launch + one synchronize:
    // no cudaMemcpy in either direction.
    const size_t head = 64;   // bytes
    int rc = f->pin.ensure(head + (size_t)L * 4 * sizeof(double));
    if (rc != STB_OK) return rc;
    double *h = (double *)f->pin.p;
    void *dptr = nullptr;
    CU(cudaHostGetDevicePointer(&dptr, h, 0));
    double *d = (double *)dptr;
    h[0] = x[0]; h[1] = y[0]; h[2] = kx[0]; h[3] = ky[0];
    volatile int32_t *hn = (volatile int32_t *)(h + 4);
    *hn = -1;

    stb_batch_t b{};
    b.struct_size = sizeof(b);
    b.nrays = 1; b.L = L;
    b.x0 = d; b.y0 = d + 1; b.kx0 = d + 2; b.ky0 = d + 3;
    b.st = st; b.dx = dx; b.dy = dy;
    b.traj = d + head / sizeof(double);
    b.nsteps = (int32_t *)(d + 4);
    KParams p{};
    bool divc = false;
    rc = fill_params(f, &b, p, divc);
    if (rc != STB_OK) return rc;
    rc = launch_trace(f, p, divc, /*refill=*/false, /*tiled=*/false, /*hoist=*/false, /*ymajor=*/false, 0);
    if (rc != STB_OK) return rc;
    CU(cudaStreamSynchronize(0));
    const int32_t n = *hn;
    if (n < 0 || n > L - 1) return fail(STB_ERR_CUDA, "stb_trace: kernel reported %d steps", (int)n);
    // scatter the interleaved states 1..n into the caller's four arrays
    // (sidtracemodule.c:41-45 writes xx[i+1], yy[i+1], kxx[i+1], kyy[i+1]); data marshalling only
    const double *t = h + head / sizeof(double);
    for (int32_t j = 1; j <= n; ++j) {
        x[j] = t[4 * j + 0]; y[j] = t[4 * j + 1]; kx[j] = t[4 * j + 2]; ky[j] = t[4 * j + 3];
    }
    *nsteps = n;
    return STB_OK;
}

int stb_field_pages(const double *xout, int64_t n, int64_t *offsets, int64_t *members, int64_t *npages)
{
    // Host helper, no device work: the page construction of Beam.field_r (beam.py:374-398), kept
    // native because folded beams of 10^4..10^6 rays make it O(rays x pages).
    if (!xout || !offsets || !members || !npages || n < 0) return fail(STB_ERR_ARG, "stb_field_pages: bad argument");
    *npages = 0;
    offsets[0] = 0;
    if (n == 0) return STB_OK;
    std::vector<char> used((size_t)n, 0);
    std::vector<std::vector<int64_t>> pages;
    pages.push_back({0});
    used[0] = 1;
    for (size_t cur = 0; cur < pages.size(); ++cur) {
        const int64_t start = pages[cur].back() + 1;
        for (int64_t j = start; j < n; ++j) {
            if (xout[j] > xout[pages[cur].back()]) {
                if (!used[(size_t)j]) { pages[cur].push_back(j); used[(size_t)j] = 1; }
            } else if (cur + 1 == pages.size()) {
                pages.push_back({j});
                used[(size_t)j] = 1;
            }
        }
    }
    int64_t np = 0, pos = 0;
    for (const auto &pg : pages) {
        if (pg.size() < 2) continue;                 // single-ray pages are dropped (beam.py:392-398)
        for (int64_t v : pg) members[pos++] = v;
        offsets[++np] = pos;
    }
    *npages = np;
    return STB_OK;
}

int stb_beam_launch(const stb_beam_t *a, void *stream)
{
    if (!a) return fail(STB_ERR_ARG, "stb_beam_launch: NULL argument");
    if (a->struct_size != sizeof(stb_beam_t))
        return fail(STB_ERR_ARG, "stb_beam_launch: struct_size %zu != %zu (ABI mismatch)", a->struct_size, sizeof(stb_beam_t));
    if (a->n < 0 || a->nfront < 2 || a->nw < 2 || a->nt < 8)
        return fail(STB_ERR_ARG, "stb_beam_launch: bad sizes");
    const void *ptrs[] = {a->lax, a->wf_x, a->t, a->c, a->x, a->y, a->kx, a->ky, a->amp};
    for (const void *q : ptrs) if (!q) return fail(STB_ERR_ARG, "stb_beam_launch: NULL pointer");
    if (a->n == 0) return STB_OK;
    // m_fine = lcm(n, nfront)
    long long g = a->n, h = a->nfront;
    while (h) { const long long tmp = g % h; g = h; h = tmp; }
    stb::BeamLaunchParams p{};
    p.n = a->n;
    p.m_fine = a->n / g * (long long)a->nfront;
    p.stride = p.m_fine / a->n;
    p.lax = a->lax; p.wfx = a->wf_x; p.nw = a->nw;
    p.t = a->t; p.c = a->c; p.nt = a->nt;
    p.bx0 = a->x0; p.by0 = a->y0; p.bkx = a->kx0; p.bky = a->ky0; p.focus = a->focus;
    p.w0 = a->waist; p.y_r = a->rayleigh;
    p.x = a->x; p.y = a->y; p.kx = a->kx; p.ky = a->ky; p.amp = a->amp;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (a->n + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    stb::beam_launch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    g_launches++;
    CU(cudaGetLastError());
    return STB_OK;
}

int stb_launch_states(const stb_launch_t *a, void *stream)
{
    if (!a) return fail(STB_ERR_ARG, "stb_launch_states: NULL argument");
    if (a->struct_size != sizeof(stb_launch_t))
        return fail(STB_ERR_ARG, "stb_launch_states: struct_size %zu != %zu (ABI mismatch)", a->struct_size, sizeof(stb_launch_t));
    if (a->n < 0 || a->nx < 2 || a->ny < 2 || a->ntx < 8 || a->nty < 8)
        return fail(STB_ERR_ARG, "stb_launch_states: bad sizes (n=%lld nx=%d ny=%d ntx=%d nty=%d)", (long long)a->n, a->nx, a->ny, a->ntx, a->nty);
    const void *ptrs[] = {a->x0_phys, a->y0_phys, a->kx_dir, a->ky_dir, a->x_nodes, a->y_nodes, a->tx, a->ty, a->coef,
                          a->x0, a->y0, a->kx0, a->ky0};
    for (const void *q : ptrs) if (!q) return fail(STB_ERR_ARG, "stb_launch_states: NULL pointer");
    if (a->n == 0) return STB_OK;
    stb::LaunchParams p{};
    p.n = a->n;
    p.x0p = a->x0_phys; p.y0p = a->y0_phys; p.kxd = a->kx_dir; p.kyd = a->ky_dir;
    p.xn = a->x_nodes; p.yn = a->y_nodes; p.nx = a->nx; p.ny = a->ny;
    p.tx = a->tx; p.ty = a->ty; p.coef = a->coef; p.ntx = a->ntx; p.nty = a->nty;
    p.w = a->w; p.c = a->c; p.nc = a->n_crit;
    p.xg = a->x0; p.yg = a->y0; p.kx = a->kx0; p.ky = a->ky0;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (a->n + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    stb::launch_states_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    g_launches++;
    CU(cudaGetLastError());
    return STB_OK;
}

int stb_pack_traj(const double *traj, int tile, int64_t ray_stride, int64_t L, int64_t nrays,
                  const int32_t *nsteps, const int64_t *offsets, double *out, void *stream)
{
    if (!traj || !nsteps || !offsets || !out) return fail(STB_ERR_ARG, "stb_pack_traj: NULL argument");
    if (tile != 0 && tile != 32 && tile != 64 && tile != 128 && tile != 256)
        return fail(STB_ERR_ARG, "stb_pack_traj: tile must be 0, 32, 64, 128 or 256");
    if (nrays <= 0) return STB_OK;
    if (!tile && ray_stride == 0) ray_stride = 4 * L;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (nrays * 32 + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    stb::pack_traj_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        traj, tile, (long long)ray_stride, (long long)L, (long long)nrays, nsteps, (const long long *)offsets, out);
    g_launches++;
    CU(cudaGetLastError());
    return STB_OK;
}

int stb_pack_traj_plan(const double *traj, int tile, const int32_t *slot_of, const int64_t *tile_off, int64_t nrays,
                       const int32_t *nsteps, const int64_t *offsets, double *out, void *stream)
{
    if (!traj || !slot_of || !tile_off || !nsteps || !offsets || !out) return fail(STB_ERR_ARG, "stb_pack_traj_plan: NULL argument");
    if (tile != 32 && tile != 64 && tile != 128 && tile != 256)
        return fail(STB_ERR_ARG, "stb_pack_traj_plan: tile must be 32, 64, 128 or 256");
    if (nrays <= 0) return STB_OK;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (nrays * 32 + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    stb::pack_plan_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        traj, tile, slot_of, (const long long *)tile_off, (long long)nrays, nsteps, (const long long *)offsets, out);
    g_launches++;
    CU(cudaGetLastError());
    return STB_OK;
}

int stb_measure_fp64_peak(double *tflops)
{
    if (!tflops) return fail(STB_ERR_ARG, "tflops is NULL");
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out = nullptr;
    const int threads = 256, blocks = sms * 8, iters = 4096;
    CU(cudaMalloc(&out, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        stb::fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        g_launches++;
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 32.0 * (double)iters * threads * (double)blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return STB_OK;
}

int stb_measure_store_bw(size_t bytes, double *gbs)
{
    if (!gbs || bytes < 32) return fail(STB_ERR_ARG, "stb_measure_store_bw: bad argument");
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out = nullptr;
    CU(cudaMalloc(&out, bytes));
    const long long nvec = (long long)(bytes / 32);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        stb::store_bw_kernel<<<sms * 16, 256>>>(out, nvec);
        g_launches++;
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double g = (double)nvec * 32.0 / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *gbs = best;
    return STB_OK;
}

int stb_measure_store_pattern(int64_t ntiles, int32_t tile, int64_t rows_per_tile, int32_t steps, int32_t chain, double *gbs)
{
    if (!gbs || ntiles < 1 || rows_per_tile < steps || steps < 1 || chain < 0 ||
        (tile != 32 && tile != 64 && tile != 128 && tile != 256))
        return fail(STB_ERR_ARG, "stb_measure_store_pattern: bad argument");
    double *out = nullptr;
    CU(cudaMalloc(&out, (size_t)ntiles * (size_t)rows_per_tile * (size_t)tile * 32));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        stb::store_pattern_kernel<<<(unsigned)ntiles, tile>>>(out, rows_per_tile, steps, chain, 1.0000001, 1e-9);
        g_launches++;
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double g = (double)ntiles * tile * steps * 32.0 / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *gbs = best;
    return STB_OK;
}

}  // extern "C"
