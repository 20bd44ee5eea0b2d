// rk_b200.cu -- C ABI of the adaptive-RK physics mode (include/sidtrace_b200_rk.h) over
// rk_kernels.cuh.  Linked into the same libsidtrace_b200.so as sidtrace_b200.cu.
#include "../../include/sidtrace_b200_rk.h"
#include "rk_kernels.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

extern "C" int stb_internal_fail(int code, const char *msg);
extern "C" void stb_internal_count_launch(void);

namespace {

constexpr int kCounters = 64;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return stb_internal_fail(code, buf);
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? STB_ERR_NOMEM : STB_ERR_CUDA,  \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

}  // namespace

struct stb_medium_s {
    int device = 0;
    int num_sms = 0;
    stbrk::Medium m{};
    int mode = 0, interp = 2;
    double *X = nullptr, *alpha = nullptr, *Y2 = nullptr;
    double2 *XA = nullptr;                   // {X, alpha} interleaved (only with an absorption table)
    double2 *XY = nullptr;                   // {X, Y^2} interleaved (tabulated-field X mode)
    unsigned long long *counter = nullptr;   // kCounters queue heads used round-robin; a slot is handed out again
                                             // only after the launch that used it last has finished (counter_ev)
    cudaEvent_t counter_ev[64] = {};
    std::atomic<unsigned> launch_seq{0};   // host threads may launch concurrently on their own streams
    std::mutex mu;                           // guards counter_ev
};

namespace {

template <int INTERP, int MODE, bool ABSORB>
int launch_rk(const stbrk::Params &p, int num_sms, cudaStream_t s, cudaEvent_t ev)
{
    auto kern = stbrk::rk_kernel<INTERP, MODE, ABSORB>;
    const size_t smem = (ABSORB && p.hist) ? (size_t)p.nbins * sizeof(double) : 0;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, stbrk::kBlock, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)num_sms * per_sm;
    const long long blocks_all = (p.nrays + stbrk::kBlock - 1) / stbrk::kBlock;
    if (blocks > blocks_all) blocks = blocks_all;
    CU(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), s));
    kern<<<(unsigned)blocks, stbrk::kBlock, smem, s>>>(p);
    stb_internal_count_launch();
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev, s));
    return STB_OK;
}

template <int INTERP>
int launch_mode(const stbrk::Params &p, int mode, bool absorb, int num_sms, cudaStream_t s, cudaEvent_t ev)
{
    if (mode == STB_RK_MODE_X_TABLE) return absorb ? launch_rk<INTERP, 2, true>(p, num_sms, s, ev) : launch_rk<INTERP, 2, false>(p, num_sms, s, ev);
    if (mode == STB_RK_MODE_X) return absorb ? launch_rk<INTERP, 1, true>(p, num_sms, s, ev) : launch_rk<INTERP, 1, false>(p, num_sms, s, ev);
    return absorb ? launch_rk<INTERP, 0, true>(p, num_sms, s, ev) : launch_rk<INTERP, 0, false>(p, num_sms, s, ev);
}

int trace_rk_host_io(stb_medium_t m, const stb_rk_t *rk, cudaStream_t s);

}  // namespace

extern "C" {

int stb_medium_create(const stb_medium_desc_t *d, stb_medium_t *out)
{
    if (!d || !out) return fail(STB_ERR_ARG, "stb_medium_create: NULL argument");
    if (d->struct_size != sizeof(stb_medium_desc_t))
        return fail(STB_ERR_ARG, "stb_medium_create: struct_size %zu != %zu (header/library mismatch)", d->struct_size,
                    sizeof(stb_medium_desc_t));
    if (!d->X) return fail(STB_ERR_ARG, "stb_medium_create: X is NULL");
    if (d->nx < 2 || d->ny < 2) return fail(STB_ERR_ARG, "stb_medium_create: nx, ny must be >= 2 (got %d, %d)", d->nx, d->ny);
    if ((long long)d->nx * d->ny > 0x7fffffffLL) return fail(STB_ERR_RANGE, "stb_medium_create: nx*ny exceeds int32");
    if (!(d->dx > 0.0) || !(d->dy > 0.0) || !std::isfinite(d->dx) || !std::isfinite(d->dy))
        return fail(STB_ERR_ARG, "stb_medium_create: dx, dy must be positive and finite");
    if (d->mode != STB_RK_MODE_O && d->mode != STB_RK_MODE_X && d->mode != STB_RK_MODE_X_TABLE)
        return fail(STB_ERR_ARG, "stb_medium_create: unknown mode %d", d->mode);
    if (d->mode == STB_RK_MODE_X && !(d->R0 > 0.0)) return fail(STB_ERR_ARG, "stb_medium_create: the X mode needs R0 > 0");
    const stb_equilibrium_t *eq = d->mode == STB_RK_MODE_X_TABLE && !d->Y2 ? d->equilibrium : nullptr;
    if (d->mode == STB_RK_MODE_X_TABLE && !d->Y2 && !eq)
        return fail(STB_ERR_ARG, "stb_medium_create: STB_RK_MODE_X_TABLE needs the Y2 node table or an equilibrium");
    if (eq) {
        if (eq->struct_size != sizeof(stb_equilibrium_t))
            return fail(STB_ERR_ARG, "stb_medium_create: equilibrium struct_size %zu != %zu", eq->struct_size, sizeof(stb_equilibrium_t));
        if (eq->nr < 2 || eq->nz < 2 || eq->nf < 2 || !eq->psi || !eq->fpol)
            return fail(STB_ERR_ARG, "stb_medium_create: the equilibrium needs psi[nr >= 2][nz >= 2] and fpol[nf >= 2]");
        if (!(eq->rdim > 0.0) || !(eq->zdim > 0.0) || !(eq->rleft > 0.0) || !(eq->sibdry != eq->simagx) || !std::isfinite(eq->y_per_tesla))
            return fail(STB_ERR_ARG, "stb_medium_create: the equilibrium needs rleft, rdim, zdim > 0, sibdry != simagx, finite y_per_tesla");
        if (!((eq->r0_cm + d->xorg) * 0.01 > 0.0))
            return fail(STB_ERR_ARG, "stb_medium_create: the medium's grid reaches R <= 0 (r0_cm = %g, xorg = %g)", eq->r0_cm, d->xorg);
    }

    if (d->interp != STB_RK_INTERP_BILINEAR && d->interp != STB_RK_INTERP_CATMULL_ROM && d->interp != STB_RK_INTERP_BSPLINE)
        return fail(STB_ERR_ARG, "stb_medium_create: unknown interpolation %d", d->interp);

    stb_medium_s *m = new (std::nothrow) stb_medium_s;
    if (!m) return fail(STB_ERR_NOMEM, "stb_medium_create: host allocation failed");
    cudaError_t e = cudaGetDevice(&m->device);
    const int nx = d->nx, ny = d->ny;
    const bool spline = d->interp == STB_RK_INTERP_BSPLINE;
    const size_t bytes = (size_t)nx * ny * sizeof(double);
    const size_t tbytes = spline ? (size_t)(nx + 2) * (ny + 2) * sizeof(double) : bytes;
    const cudaMemcpyKind kind = d->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    double *nodes = nullptr, *half = nullptr, *piv = nullptr;   // B-spline prefilter scratch
    double *y2nodes = nullptr, *eq_psi = nullptr, *eq_f = nullptr;
    do {
        if (e != cudaSuccess) break;
        cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device);
        if ((e = cudaMalloc(&m->X, tbytes)) != cudaSuccess) break;
        if (d->alpha && (e = cudaMalloc(&m->alpha, tbytes)) != cudaSuccess) break;
        if (d->mode == STB_RK_MODE_X_TABLE && (e = cudaMalloc(&m->Y2, tbytes)) != cudaSuccess) break;
        if (d->mode == STB_RK_MODE_X_TABLE) {
            // the Y^2 nodes: the caller's table, or |B|^2 of the tabulated equilibrium evaluated on the device
            if ((e = cudaMalloc(&y2nodes, bytes)) != cudaSuccess) break;
            if (d->Y2) {
                if ((e = cudaMemcpy(y2nodes, d->Y2, bytes, kind)) != cudaSuccess) break;
            } else {
                const size_t pb = (size_t)eq->nr * eq->nz * sizeof(double), fb = (size_t)eq->nf * sizeof(double);
                if ((e = cudaMalloc(&eq_psi, pb)) != cudaSuccess) break;
                if ((e = cudaMalloc(&eq_f, fb)) != cudaSuccess) break;
                if ((e = cudaMemcpy(eq_psi, eq->psi, pb, cudaMemcpyHostToDevice)) != cudaSuccess) break;
                if ((e = cudaMemcpy(eq_f, eq->fpol, fb, cudaMemcpyHostToDevice)) != cudaSuccess) break;
                stbrk::EqTable t{};
                t.psi = eq_psi; t.fpol = eq_f; t.nr = eq->nr; t.nz = eq->nz; t.nf = eq->nf;
                t.rleft = eq->rleft; t.zlow = eq->zmid - eq->zdim / 2;
                t.rdr = (double)(eq->nr - 1) / eq->rdim; t.rdz = (double)(eq->nz - 1) / eq->zdim;
                t.simagx = eq->simagx; t.rdpsi = (double)(eq->nf - 1) / (eq->sibdry - eq->simagx);
                t.r0_cm = eq->r0_cm; t.z0_cm = eq->z0_cm; t.ypt = eq->y_per_tesla;
                stbrk::equilibrium_y2_kernel<<<148 * 4, 256>>>(t, nx, ny, d->xorg, d->yorg, d->dx, d->dy, y2nodes);
                stb_internal_count_launch();
                if ((e = cudaGetLastError()) != cudaSuccess) break;
            }
        }
        if ((e = cudaMalloc(&m->counter, kCounters * sizeof(unsigned long long))) != cudaSuccess) break;
        if (!spline) {
            if ((e = cudaMemcpy(m->X, d->X, bytes, kind)) != cudaSuccess) break;
            if (d->alpha && (e = cudaMemcpy(m->alpha, d->alpha, bytes, kind)) != cudaSuccess) break;
            if (m->Y2 && (e = cudaMemcpy(m->Y2, y2nodes, bytes, cudaMemcpyDeviceToDevice)) != cudaSuccess) break;
            e = cudaDeviceSynchronize();
            break;
        }
        // reciprocal pivots of the tridiagonal (1, 4, 1) system: data-independent
        const int nmax = nx > ny ? nx : ny;
        std::vector<double> R((size_t)nmax + 1, 0.0);
        R[1] = 0.25;
        for (int i = 2; i <= nmax; ++i) R[i] = 1.0 / (4.0 - R[i - 1]);
        if ((e = cudaMalloc(&piv, R.size() * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMemcpy(piv, R.data(), R.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMalloc(&nodes, bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&half, (size_t)(nx + 2) * ny * sizeof(double))) != cudaSuccess) break;
        for (int which = 0; which < 3 && e == cudaSuccess; ++which) {
            const double *src = which == 2 ? y2nodes : which ? d->alpha : d->X;
            double *dst = which == 2 ? m->Y2 : which ? m->alpha : m->X;
            if (!src) continue;
            if ((e = cudaMemcpy(nodes, src, bytes, which == 2 ? cudaMemcpyDeviceToDevice : kind)) != cudaSuccess) break;
            // rows: [ny][nx] -> [ny][nx+2]; columns: [ny][nx+2] -> [ny+2][nx+2]
            stbrk::bspline_prefilter_kernel<<<(ny + 127) / 128, 128>>>(nodes, nx, 1, half, nx + 2, 1, piv, nx, ny);
            stbrk::bspline_prefilter_kernel<<<(nx + 2 + 127) / 128, 128>>>(half, 1, nx + 2, dst, 1, nx + 2, piv, ny, nx + 2);
            stb_internal_count_launch(); stb_internal_count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    } while (0);
    if (e == cudaSuccess && m->alpha) {
        const long long nt = (long long)(tbytes / sizeof(double));
        if ((e = cudaMalloc(&m->XA, (size_t)nt * sizeof(double2))) == cudaSuccess) {
            stbrk::interleave2_kernel<<<148 * 8, 256>>>(m->X, m->alpha, m->XA, nt);
            stb_internal_count_launch();
            if ((e = cudaGetLastError()) == cudaSuccess) e = cudaDeviceSynchronize();
        }
    }
    if (e == cudaSuccess && m->Y2) {     // also with absorption: the end-of-ray Hamiltonian is evaluated without the alpha table
        const long long nt = (long long)(tbytes / sizeof(double));
        if ((e = cudaMalloc(&m->XY, (size_t)nt * sizeof(double2))) == cudaSuccess) {
            stbrk::interleave2_kernel<<<148 * 8, 256>>>(m->X, m->Y2, m->XY, nt);
            stb_internal_count_launch();
            if ((e = cudaGetLastError()) == cudaSuccess) e = cudaDeviceSynchronize();
        }
    }
    cudaFree(nodes); cudaFree(half); cudaFree(piv); cudaFree(y2nodes); cudaFree(eq_psi); cudaFree(eq_f);
    if (e != cudaSuccess) {
        const int rc = fail(e == cudaErrorMemoryAllocation ? STB_ERR_NOMEM : STB_ERR_CUDA, "stb_medium_create: %s",
                            cudaGetErrorString(e));
        cudaFree(m->X); cudaFree(m->alpha); cudaFree(m->XA); cudaFree(m->XY); cudaFree(m->Y2); cudaFree(m->counter);
        delete m;
        return rc;
    }
    m->mode = d->mode;
    m->interp = d->interp;
    m->m.X = m->X; m->m.alpha = m->alpha; m->m.XA = m->XA; m->m.Y2 = m->Y2; m->m.XY = m->XY;
    m->m.nx = d->nx; m->m.ny = d->ny;
    m->m.xorg = d->xorg; m->m.yorg = d->yorg;
    m->m.rdx = 1.0 / d->dx; m->m.rdy = 1.0 / d->dy;
    m->m.Y0R0 = d->Y0 * d->R0; m->m.R0 = d->R0;
    *out = m;
    return STB_OK;
}

int stb_medium_table(stb_medium_t m, int which, double *host_out, size_t count)
{
    if (!m || !host_out) return fail(STB_ERR_ARG, "stb_medium_table: NULL argument");
    if (which < 0 || which > 2) return fail(STB_ERR_ARG, "stb_medium_table: which must be 0 (X), 1 (alpha) or 2 (Y^2)");
    const double *src = which == 2 ? m->Y2 : which ? m->alpha : m->X;
    if (!src) return fail(STB_ERR_ARG, "stb_medium_table: the medium has no %s table", which == 2 ? "Y^2" : "absorption");
    const int pad = m->interp == STB_RK_INTERP_BSPLINE ? 2 : 0;
    const size_t have = (size_t)(m->m.nx + pad) * (size_t)(m->m.ny + pad);
    if (count != have) return fail(STB_ERR_ARG, "stb_medium_table: the table holds %zu doubles, not %zu", have, count);
    CU(cudaMemcpy(host_out, src, have * sizeof(double), cudaMemcpyDeviceToHost));
    return STB_OK;
}

int stb_medium_destroy(stb_medium_t m)
{
    if (!m) return STB_OK;
    cudaFree(m->X); cudaFree(m->alpha); cudaFree(m->XA); cudaFree(m->XY); cudaFree(m->Y2); cudaFree(m->counter);
    for (auto &e : m->counter_ev) if (e) cudaEventDestroy(e);
    delete m;
    return STB_OK;
}

int stb_trace_rk(stb_medium_t m, const stb_rk_t *rk, void *stream)
{
    if (!m || !rk) return fail(STB_ERR_ARG, "stb_trace_rk: NULL argument");
    if (rk->struct_size != sizeof(stb_rk_t))
        return fail(STB_ERR_ARG, "stb_trace_rk: struct_size %zu != %zu (header/library mismatch)", rk->struct_size, sizeof(stb_rk_t));
    if (rk->nrays < 0) return fail(STB_ERR_ARG, "stb_trace_rk: nrays < 0");
    if (rk->nrays == 0) return STB_OK;
    if (!rk->x0 || !rk->y0 || !rk->px0 || !rk->py0) return fail(STB_ERR_ARG, "stb_trace_rk: launch state pointers must not be NULL");
    if (!(rk->rtol >= 0.0) || !(rk->atol_r > 0.0) || !(rk->atol_p > 0.0))
        return fail(STB_ERR_ARG, "stb_trace_rk: need rtol >= 0 and atol_r, atol_p > 0");
    if (!(rk->h0 > 0.0) || !(rk->hmin > 0.0) || !(rk->hmax >= rk->h0) || !(rk->h0 >= rk->hmin))
        return fail(STB_ERR_ARG, "stb_trace_rk: need 0 < hmin <= h0 <= hmax");
    if (!(rk->tau_max > 0.0)) return fail(STB_ERR_ARG, "stb_trace_rk: tau_max must be positive");
    if (rk->refill_min < 0 || rk->refill_min > 32) return fail(STB_ERR_ARG, "stb_trace_rk: refill_min must be in [0, 32]");
    if (rk->stop_axis < 0 || rk->stop_axis > 2) return fail(STB_ERR_ARG, "stb_trace_rk: stop_axis must be 0, 1 or 2");
    if (rk->stop_axis && !(rk->stop_tol > 0.0)) return fail(STB_ERR_ARG, "stb_trace_rk: stop_tol must be positive");
    if (rk->traj && rk->L < 1) return fail(STB_ERR_ARG, "stb_trace_rk: traj needs L >= 1");
    if (rk->traj_aux && !rk->traj) return fail(STB_ERR_ARG, "stb_trace_rk: traj_aux needs traj");
    if (rk->hist) {
        if (!m->alpha) return fail(STB_ERR_ARG, "stb_trace_rk: the deposition histogram needs a medium with an absorption table");
        if (rk->nbins < 1 || rk->nbins > 28000) return fail(STB_ERR_ARG, "stb_trace_rk: nbins must be in [1, 28000]");
        if (!(rk->r_lcs > 0.0) || !(rk->rho_max > 0.0)) return fail(STB_ERR_ARG, "stb_trace_rk: r_lcs and rho_max must be positive");
    }
    int dev = -1;
    CU(cudaGetDevice(&dev));
    if (dev != m->device)
        return fail(STB_ERR_ARG, "stb_trace_rk: the medium lives on device %d but the current device is %d", m->device, dev);

    if (rk->flags & STB_RK_F_HOST_IO) return trace_rk_host_io(m, rk, (cudaStream_t)stream);

    stbrk::Params p{};
    p.m = m->m;
    p.nrays = rk->nrays;
    p.x0 = rk->x0; p.y0 = rk->y0; p.px0 = rk->px0; p.py0 = rk->py0;
    p.xscale = rk->xscale; p.yscale = rk->yscale; p.amp = rk->amp;
    p.rtol = rk->rtol; p.atol_r = rk->atol_r; p.atol_p = rk->atol_p;
    p.h0 = rk->h0; p.hmin = rk->hmin; p.hmax = rk->hmax; p.tau_max = rk->tau_max;
    p.max_steps = rk->max_steps;
    p.refill_min = (rk->flags & STB_RK_F_LANE_REFILL) ? 1 : (rk->refill_min > 0 ? rk->refill_min : 24);
    p.stop_axis = rk->stop_axis; p.stop_val = rk->stop_val; p.stop_tol = rk->stop_tol;
    p.end_state = rk->end_state; p.tau_end = rk->tau_end; p.phi_end = rk->phi_end; p.sig_end = rk->sig_end;
    p.n2_min = rk->n2_min; p.ham_end = rk->ham_end;
    p.nacc = rk->nacc; p.nrej = rk->nrej; p.status = rk->status;
    p.turn_index = rk->turn_index; p.turn_state = rk->turn_state;
    p.traj = rk->traj; p.traj_aux = rk->traj_aux; p.L = rk->L;
    p.hist = rk->hist; p.nbins = rk->nbins;
    p.xaxis = rk->xaxis; p.yaxis = rk->yaxis;
    if (rk->hist) {
        p.r_rlcs = 1.0 / rk->r_lcs;
        p.bins_per_rho = (double)rk->nbins / rk->rho_max;
    }
    std::lock_guard<std::mutex> lk(m->mu);
    const unsigned slot = m->launch_seq.fetch_add(1) % kCounters;
    p.counter = m->counter + slot;
    if (m->counter_ev[slot]) CU(cudaEventSynchronize(m->counter_ev[slot]));
    else CU(cudaEventCreateWithFlags(&m->counter_ev[slot], cudaEventDisableTiming));
    cudaEvent_t ev = m->counter_ev[slot];

    cudaStream_t s = (cudaStream_t)stream;
    const bool absorb = m->alpha != nullptr;
    switch (m->interp) {
        case STB_RK_INTERP_BILINEAR: return launch_mode<0>(p, m->mode, absorb, m->num_sms, s, ev);
        case STB_RK_INTERP_CATMULL_ROM: return launch_mode<1>(p, m->mode, absorb, m->num_sms, s, ev);
        default: return launch_mode<2>(p, m->mode, absorb, m->num_sms, s, ev);
    }
}

}  // extern "C"

namespace {

// STB_RK_F_HOST_IO: one device arena for every array of the call, copies in, the ordinary device
// launch, copies out.  Synchronous; meant for non-CUDA hosts and small batches.
int trace_rk_host_io(stb_medium_t m, const stb_rk_t *rk, cudaStream_t s)
{
    const size_t n = (size_t)rk->nrays;
    struct Slot { const void *host_in; void *host_out; size_t bytes; size_t off; };
    std::vector<Slot> slots;
    size_t total = 0;
    auto add = [&](const void *in, void *out, size_t bytes) -> size_t {
        if (!in && !out) return (size_t)-1;
        const size_t off = total;
        slots.push_back({in, out, bytes, off});
        total += (bytes + 255) & ~(size_t)255;
        return off;
    };
    const size_t d8 = sizeof(double), i4 = sizeof(int32_t);
    const size_t o_x0 = add(rk->x0, nullptr, n * d8), o_y0 = add(rk->y0, nullptr, n * d8);
    const size_t o_px = add(rk->px0, nullptr, n * d8), o_py = add(rk->py0, nullptr, n * d8);
    const size_t o_xs = add(rk->xscale, nullptr, n * d8), o_ys = add(rk->yscale, nullptr, n * d8);
    const size_t o_amp = add(rk->amp, nullptr, n * d8);
    const size_t o_end = add(nullptr, rk->end_state, 4 * n * d8);
    const size_t o_tau = add(nullptr, rk->tau_end, n * d8), o_phi = add(nullptr, rk->phi_end, n * d8);
    const size_t o_sig = add(nullptr, rk->sig_end, n * d8), o_n2 = add(nullptr, rk->n2_min, n * d8);
    const size_t o_ham = add(nullptr, rk->ham_end, n * d8);
    const size_t o_nacc = add(nullptr, rk->nacc, n * i4), o_nrej = add(nullptr, rk->nrej, n * i4);
    const size_t o_stat = add(nullptr, rk->status, n * i4);
    const size_t o_turn = add(nullptr, rk->turn_index, n * i4), o_tst = add(nullptr, rk->turn_state, 4 * n * d8);
    const size_t L = rk->traj ? (size_t)rk->L : 0;
    // trajectories and the histogram are read-modify-write for the caller (untouched tail, accumulation)
    const size_t o_traj = add(rk->traj, rk->traj, n * L * 4 * d8), o_aux = add(rk->traj_aux, rk->traj_aux, n * L * 2 * d8);
    const size_t o_hist = add(rk->hist, rk->hist, rk->hist ? (size_t)rk->nbins * d8 : 0);

    char *arena = nullptr;
    CU(cudaMalloc(&arena, total ? total : 256));
    int rc = STB_OK;
    cudaError_t e = cudaSuccess;
    for (const Slot &sl : slots)
        if (sl.host_in && sl.bytes && (e = cudaMemcpyAsync(arena + sl.off, sl.host_in, sl.bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
    if (e == cudaSuccess) {
        stb_rk_t d = *rk;
        d.flags &= ~STB_RK_F_HOST_IO;
        auto at = [&](size_t off) -> void * { return off == (size_t)-1 ? nullptr : (void *)(arena + off); };
        d.x0 = (const double *)at(o_x0); d.y0 = (const double *)at(o_y0);
        d.px0 = (const double *)at(o_px); d.py0 = (const double *)at(o_py);
        d.xscale = (const double *)at(o_xs); d.yscale = (const double *)at(o_ys); d.amp = (const double *)at(o_amp);
        d.end_state = (double *)at(o_end); d.tau_end = (double *)at(o_tau); d.phi_end = (double *)at(o_phi);
        d.sig_end = (double *)at(o_sig); d.n2_min = (double *)at(o_n2); d.ham_end = (double *)at(o_ham);
        d.nacc = (int32_t *)at(o_nacc); d.nrej = (int32_t *)at(o_nrej); d.status = (int32_t *)at(o_stat);
        d.turn_index = (int32_t *)at(o_turn); d.turn_state = (double *)at(o_tst);
        d.traj = (double *)at(o_traj); d.traj_aux = (double *)at(o_aux); d.hist = (double *)at(o_hist);
        rc = stb_trace_rk(m, &d, s);
    }
    if (e == cudaSuccess && rc == STB_OK)
        for (const Slot &sl : slots)
            if (sl.host_out && sl.bytes && (e = cudaMemcpyAsync(sl.host_out, arena + sl.off, sl.bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(arena);
    if (rc != STB_OK) return rc;
    if (e != cudaSuccess) return fail(STB_ERR_CUDA, "stb_trace_rk (host I/O): %s", cudaGetErrorString(e));
    return STB_OK;
}

}  // namespace

