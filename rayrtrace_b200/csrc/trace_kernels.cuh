// trace_kernels.cuh -- sm_100a kernels of the sidtrace hot path.
//
// One thread owns one ray at a time and steps it in a plain per-thread loop (no
// vote per step).  Three schedules hand rays to threads:
//   static       one thread per ray; the hardware CTA scheduler balances the load;
//   tile refill  persistent warps claim 32 consecutive rays at a time (TILED);
//   lane refill  persistent warps vote (ballot) once every few steps on which lanes
//                still own a live ray; lanes whose ray left the grid or hit the cap
//                have retired it (written the per-ray results), and the warp claims
//                replacements from a global queue with ONE aggregated atomic, handed
//                out by lane rank -- a reflected or trapped ray (up to MaxStep =
//                100000 steps, ray.py:11) never holds 31 idle lanes hostage.
//
// Arithmetic contract (bit-exact with sidtracemodule.c:36-48 as gcc -O2
// compiles it: no FMA, left-to-right):
//     t = ((st*st)*M)/2 ; u = k*st ; v = t+u ; w = v/d ; x' = x + w
//     k' = k + st*M
// Every FP64 operation below is an explicit round-to-nearest intrinsic so no
// contraction can happen whatever the compiler flags are.
//
// The two divisions per step have a loop-invariant divisor (dx, dy).  With
// rd = RN(1/d) computed once on the host, the sequence
//     q = RN(v*rd) ; r = RN(v - q*d) (exact, one FMA) ; q' = RN(q + r*rd)
// returns the correctly rounded quotient RN(v/d) (Markstein's correction step:
// q is within one ulp of v/d, so r is exactly representable and the final FMA
// rounds the exact quotient) provided nothing under- or overflows.  div_const()
// guards the exponent range and falls back to the generic IEEE division
// otherwise, so the result is bit-identical to `v / d` for every input.
// tests/test_gpu_parity.py checks the two paths against each other and against
// the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stb {

constexpr int kBlock = 128;   // threads per CTA (4 warps); the tiled trajectory kernel also runs with 256
#ifndef STB_MINB_CROSS
#define STB_MINB_CROSS 5      // crossing / lookup passes, static grid: 96 registers, 20 warps/SM (1 / 4 / 5 / 6 CTAs per SM measured:
                              // cross 4.48 / 4.51 / 4.17 / 4.47 ms, lookup 9.99 / 6.94 / 6.64 / 6.81 ms per 10^6 rays,
                              // profiles/r02/kernel_sweep_restructured.txt)
#endif
#ifndef STB_MINB_CROSS_REFILL
#define STB_MINB_CROSS_REFILL 7   // the same passes on the persistent schedule (what Beam / ft2.phases launch): 72 registers with
                                  // a few spills, 28 warps/SM.  5 / 6 / 7 / 8 CTAs per SM: cross 4.11 / 3.96 / 3.75 / 3.85 ms,
                                  // lookup 5.83 / 5.62 / 5.37 / 7.36 ms (profiles/r02/cross_occupancy.txt); the static grid
                                  // loses with more than 5
#endif
#ifndef STB_MINB_EXTRAS
#define STB_MINB_EXTRAS 8     // fused sums / histogram: 64 registers without spills, 32 warps/SM: 1-4 % over the 72 registers ptxas takes
#endif                        // when left alone -- except for the fused-sum pass on the persistent schedule (-8 %), which keeps 1
                              // (profiles/r02/extras_occupancy.txt)
#ifndef STB_MINB
#define STB_MINB 12           // min resident CTAs/SM asked of ptxas for the lean variants: 40 registers,
#endif                        // 48 warps/SM (measured: +4..7 % in endpoint mode, neutral for trajectories;
                              // 16 -> 32 registers spills and loses 20 %, profiles/r01/occupancy_layout_sweep.txt)

struct KParams {
    const double2 *field;   // {ax, ay} per cell; cell (iy, ix) sits at iy*csy + ix*csx
    const double4 *field4;  // HOIST: {RN(RN(sst*ax)/2), RN(RN(sst*ay)/2), RN(st*ax), RN(st*ay)} per cell for the
                            // launch's global step (built once per (field, st) by hoist_kernel)
    int nxm1, nym1, ncell;
    int csx, csy;                        // cell strides: (1, nxm1) = x-major rows, the reference's a[iy*(nx-1) + ix]
                                         // (sidtracemodule.c:38); (nym1, 1) = the y-major copy (STB_F_YMAJOR): a
                                         // ray that travels along y then finds its next cells in the SAME 128-byte
                                         // line instead of one row (nxm1 cells) further on
    int Lm1;                             // L - 1: the guard compares the state index against it (constant-bank operand)
    double fx, fy;                       // (double)(nx-1), (double)(ny-1)
    long long nrays;
    int L;
    const double *x0, *y0, *kx0, *ky0;
    double st, sst;                      // sst = RN(st*st), computed once on the host
    const double *st_ray;
    double dx, dy, rdx, rdy;
    double *traj;
    long long stride;                    // doubles per ray (ray-major layout)
    int tile;                            // rays per tile (tiled layout): 32, 64, 128 or 256
    // perm: slot -> ray (-1 = empty).  With tile_off: the packed tiled layout of stb_traj_plan (first row of every
    // tile, slots = ntiles * tile).  Without: just the ORDER the rays are traced in (coherent_order: launch sets in
    // random order) -- per-ray outputs and ray-major trajectories land at the ray's own index, tiled trajectories at
    // the slot
    const int *perm;
    const long long *tile_off;
    long long nslots;
    int *nsteps;
    double *endst;
    int *turn;
    double *turn_state;                  // [4][nrays]
    double *lax_end, *ph_end;
    double *hist;
    int nbins;
    const double *amp;
    double xorg, yorg, r_lcs, rho_max;
    double r_rlcs, r_rhomax;             // RN(1/r_lcs), RN(1/rho_max) for the constant-divisor quotients
    double hist_dr;                      // radial bin width in cm
    int hist_annulus;                    // bins at least a step wide: use the squared-radius annulus test
    int hist_table;                      // ... and keep every bin's (shrunk) squared-radius bounds in shared memory
    unsigned long long *counter;         // ray queue head (refill schedule)
    // receiver-line crossing (CROSS)
    int cross_axis;                      // 1: y = cross_val, 2: x = cross_val (physical cm)
    double cross_val;
    const double *x0p, *y0p;             // physical launch points
    int *cross_count, *cross_index, *cross_flags;
    double *cross_state;                 // [6][nrays]
    // equal-phase lookup (LOOKUP): two target phases per ray -> point of the ray at that phase
    const double *lookup_ph;             // [2][nrays]
    double *lookup_xy;                   // [4][nrays]  xA, yA, xB, yB (physical)
};

// Smallest non-negative root s of k*s + a*s^2/2 = val - c0 (ray.py:174-193); NaN when none.
__device__ __forceinline__ double cross_solve(double c0, double k, double a, double val, double st1)
{
    double st = st1;
    if (a != 0.0) {
        const double disc = k * k + 2.0 * a * (val - c0);
        if (!(disc >= 0.0)) return __longlong_as_double(0x7ff8000000000000LL);
        const double sq = sqrt(disc);
        const double r1 = (-k - sq) / a, r2 = (-k + sq) / a;
        const bool ok1 = r1 >= 0.0, ok2 = r2 >= 0.0;
        if (!ok1 && !ok2) return __longlong_as_double(0x7ff8000000000000LL);
        st = ok1 && ok2 ? fmin(r1, r2) : (ok1 ? r1 : r2);
        if (st == 0.0 && val != c0 && k != 0.0) st = (val - c0) / k;
    } else if (k != 0.0) {
        st = (val - c0) / k;
    }
    return st;
}

// L2 residency of the field while trajectories stream through the L2.  A 26 GB store stream evicts the
// 16 MB field: ncu saw 60 % of the field reads miss in L2 and the warps wait on them (long scoreboard,
// profiles/r01).  In the TRAJ kernels the field gathers therefore carry an L2 evict_last policy and the
// trajectory stores evict_first: 6436 -> 6810 GB/s on BASELINE config 3 (either hint alone: 6620;
// a persisting-L2 carve-out of 20 / 40 MB: 6630 / 5060; profiles/r02/l2_policy_sweep.txt).  Kernels
// without trajectory output keep plain loads (the hinted load costs them 15 %).
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// What bounds the kernels WITHOUT trajectory output.  Measured on B200 (tools/probe/latency_probe.cu,
// profiles/r02/latency_probe.txt, occupancy_curve_endpoints.txt, dual_ray_endpoint.txt): FP64 operations 8.3
// cycles, F2I.F64 ~19, a warp-uniform 32-byte gather 60-70 cycles from L1 and 300+ from L2, one ray-step of one
// warp 437 cycles in isolation and ~570 in the kernel.  Up to 8 warps per scheduler the kernel is bound by that
// latency (its run time does not move between 1 and 8); at the 12 the 40-register cap allows it is not any more:
// two rays per thread (72 instead of 48 independent chains per SM, endpoint_dual_kernel) buy 3.5 %, and forcing
// every gather onto one L1-resident cell buys 11 % more.  What is left is the instruction stream itself -- an
// FP64 instruction holds its scheduler's dispatch port for two cycles, so a ray-step costs (instructions + 14)
// dispatch cycles, and 7.2e11 ray-steps/s is 94 % of what the 592 schedulers can dispatch.
// A ray enters a new cell -- a new 32-byte sector, whatever the layout -- almost every step, and the four warps
// of a CTA ride the same cells in convoy, so each of them waits for the same L2 fill: that is the 11 %.
// Tried against it, measured, dropped (profiles/r02/prefetch_dropped.txt): `prefetch.global.L1` of the
// neighbouring cells (CCTL.E.PF1 leaves the probe's chain at 438 cycles: no effect, 2 % slower in the
// kernel); real 4-byte "touch" loads of the cells 2 ahead and behind along the fast axis of the y-major
// table (the probe's chain drops to 289 cycles and the kernel's demand-gather stall samples fall 4x, but the
// two extra LSU instructions per step cost more dispatch cycles than the wait they remove: 1.28 -> 1.40 ms,
// and 1.9x slower on rays in random order, where every lane touches a sector of its own).
constexpr int kTablePadCells = 8;   // cells of padding on either side of every cell table (room for such experiments)

// one 32-byte cell of the hoisted table = one LDG.E.ENL2.256.CONSTANT
template <bool KEEP>
__device__ __forceinline__ double4 ld_cell4(const double4 *p)
{
    double4 v;
    if (KEEP)
        asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
                     : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p), "l"(l2_policy_evict_last()));
    else
        asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// one 16-byte cell {ax, ay}
template <bool KEEP>
__device__ __forceinline__ double2 ld_cell2(const double2 *p)
{
    if (KEEP) {
        double2 v;
        asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(l2_policy_evict_last()));
        return v;
    }
    return __ldg(p);
}

// one ray-step = one 32-byte sector = one STG.E.ENL2.256
// (st.global.cs -- evict-first in L1 -- measured the same within noise in round 1; the L2 policy is what counts)
template <bool STREAM = false>
__device__ __forceinline__ void st_state(double *p, double a, double b, double c, double d)
{
    if (STREAM)
        asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;"
                     :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(l2_policy_evict_first()) : "memory");
    else
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};"
                     :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// Correctly rounded v / d for a loop-invariant divisor d with rd = RN(1/d).
template <bool DIVC>
__device__ __forceinline__ double div_const(double v, double d, double rd)
{
    if (!DIVC) return __ddiv_rn(v, d);
    const unsigned e = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
    // |v| in [2^-400, 2^400]: q, r, q' all normal, r exact (host checked |d| in [2^-400, 2^400]).
    if (e - 623u <= 800u) {
        const double q = __dmul_rn(v, rd);
        const double r = __fma_rn(-q, d, v);
        return __fma_rn(r, rd, q);
    }
    if (v == 0.0) return __dmul_rn(v, rd);   // +-0 / d = +-0 (sign of d == sign of rd)
    return __ddiv_rn(v, d);                  // subnormal, huge, inf, nan: generic path
}

// Both quotients of a step.  One combined exponent test sends the (overwhelmingly common) case
// "vx and vy both in the guarded range" down a branch-free path where the two 3-instruction
// sequences interleave; anything else (a zero, a subnormal, inf, nan) takes the per-quotient path.
template <bool DIVC>
__device__ __forceinline__ void div_pair(double vx, double vy, double dx, double dy, double rdx, double rdy,
                                         double &wx, double &wy)
{
    if (DIVC) {
        const unsigned ex = ((unsigned)__double2hiint(vx) & 0x7ff00000u) - (623u << 20);
        const unsigned ey = ((unsigned)__double2hiint(vy) & 0x7ff00000u) - (623u << 20);
        if (max(ex, ey) <= (800u << 20)) {
            const double qx = __dmul_rn(vx, rdx), qy = __dmul_rn(vy, rdy);
            const double rx = __fma_rn(-qx, dx, vx), ry = __fma_rn(-qy, dy, vy);
            wx = __fma_rn(rx, rdx, qx);
            wy = __fma_rn(ry, rdy, qy);
            return;
        }
    }
    wx = div_const<DIVC>(vx, dx, rdx);
    wy = div_const<DIVC>(vy, dy, rdy);
}

// 0 <= v < n evaluated on the integer pipe; c = __double2int_rz(v).  See the guard in step().
__device__ __forceinline__ bool in_range(double v, int c, int n)
{
    const unsigned hi = (unsigned)__double2hiint(v);
    const bool nonneg = (hi < 0x7ff00000u) | ((hi == 0x80000000u) & (__double2loint(v) == 0));
    return nonneg & ((unsigned)c < (unsigned)n);
}

__device__ __forceinline__ double kmod(double kx, double ky)
{   // ray.py:115  np.sqrt(kx**2 + ky**2)
    return __dsqrt_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)));
}

// Adds w to shared-memory bin b for every active lane; lanes of the warp that hit the same bin are combined
// first (match.any) so each distinct bin costs one shared atomic per warp-flush.  The peers' values are summed
// in log2(peers) rounds (every round each remaining peer takes the value of the next remaining one, then every
// second peer drops out) -- neighbouring rays change bin together, so the usual group is the whole warp: 5
// rounds instead of the 32 sequential shuffles of round 1's lane-order loop.  Deterministic for a given set of
// active lanes (a fixed tree).
__device__ __forceinline__ void hist_add(double *sh, int b, double w)
{
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
#ifndef STB_HIST_NO_BUTTERFLY
    if (active == 0xffffffffu) {
        // the whole warp flushes, and (neighbouring rays) usually into one bin: xor butterfly, one atomic
        int same;
        __match_all_sync(active, b, &same);
        if (same) {
            double sum = w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(active, sum, o));
            if (lane == 0 && b >= 0) atomicAdd(&sh[b], sum);
            return;
        }
    }
#endif
    unsigned peers = __match_any_sync(active, b);
    const int leader = __ffs(peers) - 1;
    int rel = __popc(peers & ((1u << lane) - 1u));          // my rank among my peers
    peers &= 0xfffffffeu << lane;                           // peers above me
    double sum = w;
    while (__any_sync(active, peers != 0)) {
        const int next = __ffs(peers);                      // 1 + lane of the next remaining peer above me, 0 = none
        const double t = __shfl_sync(active, sum, next ? next - 1 : lane);
        if (next) sum = __dadd_rn(sum, t);
        const bool done = rel & 1;                          // odd ranks have handed their value down: they drop out
        peers &= __ballot_sync(active, !done);
        rel >>= 1;
    }
    if (lane == leader && b >= 0) atomicAdd(&sh[b], sum);
}

// What a crossing capture needs besides the registers of the segment's start state.  Captures are rare
// (once per ray), so this code stays out of line: the stepping loop keeps its registers for the steps.
struct CrossCtx {
    const double2 *field;
    const double *x0p, *y0p, *xg0, *yg0;    // physical launch points, launch state (grid units)
    double *cross_state;                      // [6][nrays] or nullptr
    long long nrays;
    double dx, dy, val;
    bool axis_y;
};
struct CrossPoint { double x, y, lax, ph, s, kn; };

// Crossing point inside (or extrapolated beyond) the segment that starts at the kept state
// (sx, sy, skx, sky) in cell sind with running sums (slax, sph): ray.py:174-216.  Writes the per-ray
// crossing outputs; false when the segment has no usable root.
__device__ __noinline__ bool cross_capture(const CrossCtx c, long long ray, double sx, double sy, double skx, double sky,
                                           int sind, double slax, double sph, double st1, CrossPoint *out)
{
    const double xp = c.x0p[ray] + (sx - c.xg0[ray]) * c.dx, yp = c.y0p[ray] + (sy - c.yg0[ray]) * c.dy;   // ray.py:105-106
    const double2 pm = __ldg(&c.field[sind]);
    const double s = cross_solve(c.axis_y ? yp : xp, c.axis_y ? sky : skx, c.axis_y ? pm.y : pm.x, c.val, st1);
    if (!(s == s) || s == 0.0) return false;
    const double nkx = skx + pm.x * s, nky = sky + pm.y * s;
    const double nx_ = xp + skx * s + pm.x / 2 * s * s, ny_ = yp + sky * s + pm.y / 2 * s * s;
    const double kp = sqrt(skx * skx + sky * sky), kn = sqrt(nkx * nkx + nky * nky);
    const double kav = (kp + kn) / 2;
    const double nlax = slax + s * kav;
    const double nph = sph + (nlax - slax) * kav;
    if (c.cross_state) {
        c.cross_state[0 * c.nrays + ray] = nx_;  c.cross_state[1 * c.nrays + ray] = ny_;
        c.cross_state[2 * c.nrays + ray] = nkx;  c.cross_state[3 * c.nrays + ray] = nky;
        c.cross_state[4 * c.nrays + ray] = nlax; c.cross_state[5 * c.nrays + ray] = nph;
    }
    out->x = nx_; out->y = ny_; out->lax = nlax; out->ph = nph; out->s = s; out->kn = kn;
    return true;
}

// An existing point lies on the receiver line (ray.py:142-143): it is reported as it stands.
__device__ __noinline__ void cross_exact_hit(const CrossCtx c, long long ray, double x, double y, double kx, double ky,
                                             double lax, double ph)
{
    if (!c.cross_state) return;
    c.cross_state[0 * c.nrays + ray] = c.x0p[ray] + (x - c.xg0[ray]) * c.dx;
    c.cross_state[1 * c.nrays + ray] = c.y0p[ray] + (y - c.yg0[ray]) * c.dy;
    c.cross_state[2 * c.nrays + ray] = kx;  c.cross_state[3 * c.nrays + ray] = ky;
    c.cross_state[4 * c.nrays + ray] = lax; c.cross_state[5 * c.nrays + ray] = ph;
}

// TILED (trajectory layout [ray/T][step][ray%T][4], T = p.tile): a warp owns 32 consecutive
// rays; its lanes step in lock-step, so every warp-step is ONE fully coalesced 1 KB row store,
// the warps of a CTA write adjacent rows, and consecutive steps are consecutive rows --
// sequential DRAM pages instead of 32 scattered sectors.  Rays are claimed 32 at a time
// (a lane-granular refill would scatter a new ray's rows); per-ray results are retired when
// the whole warp is done, as coalesced rows.
//
// CROSS (needs EXTRAS): first crossing of a receiver line y = val / x = val with the semantics of
// the reference's Ray.addpoint (ray.py:131-224): exact hits, sign-change segments solved with the
// start cell's constant acceleration, or extrapolation past the last kept point.  Tolerance-level
// parity (the reference squares through libm pow there), so plain arithmetic is used.
// STRAY = false: one global step for the whole launch (the reference's case, den.py:107): st and
// st*st stay in the constant bank instead of four registers per thread.
// LOOKUP (needs CROSS): the ray as Beam.field_r sees it after addpoint -- the polyline of kept
// states with the crossing point inserted and the eikonal phase recomputed (ray.py:205-216,
// 122-126) -- is sampled at two target phases by linear interpolation / end-segment
// extrapolation, exactly what interp1d(r.ph, r.x, fill_value='extrapolate') does in
// beam.py:410-413.  Nothing is stored: the ray is simply traced again.
//
// BLOCK: threads per CTA.  128 everywhere except the static tiled trajectory kernel, which runs
// 256-thread CTAs over 256-ray tiles: a CTA-step is then one contiguous 8 KB row (+1.3 % over 4 KB
// rows, profiles/r01/kernel_sweep_block256_cs.txt; the O(rays)-output modes lose 0.5-3 % with 256).
// HOIST (needs STRAY = false): with one global step the four products of a step that depend only
// on the cell -- RN(RN(sst*Mx)/2), RN(RN(sst*My)/2), RN(st*Mx), RN(st*My), sidtracemodule.c:41-45 --
// are read from a per-(field, st) table built once by hoist_kernel with the very same operations on
// the very same operands (bit-exact by construction): one 32-byte gather per step, 14 instead of
// 20 FP64-pipe instructions.
template <bool TRAJ, bool EXTRAS, bool HIST, bool DIVC, bool REFILL, bool TILED = false, bool CROSS = false,
          bool STRAY = true, bool LOOKUP = false, int BLOCK = kBlock, bool HOIST = false>
__global__ void __launch_bounds__(BLOCK, CROSS ? (REFILL ? STB_MINB_CROSS_REFILL : STB_MINB_CROSS)
                                          : (HIST || EXTRAS) ? ((EXTRAS && !HIST && REFILL) ? 1 : STB_MINB_EXTRAS) : STB_MINB * kBlock / BLOCK)
trace_kernel(const KParams p)
{
    extern __shared__ double sh_hist[];
    // HIST with hist_table: sh_lo[b], sh_hi[b] = the squared physical radii between which a point is certainly in
    // bin b (the bin's edges shrunk by 1e-12, far more than the rounding of the exact formula) -- the same
    // expressions the per-lane annulus below uses
    double *sh_lo = sh_hist + p.nbins, *sh_hi = sh_hist + 2 * p.nbins;
    if (HIST) {
        for (int b = threadIdx.x; b < p.nbins; b += BLOCK) {
            sh_hist[b] = 0.0;
            if (p.hist_table) {
                const double a = p.hist_dr * (double)b, b2 = p.hist_dr * (double)(b + 1);
                sh_lo[b] = a * a * (1.0 + 1e-12);
                sh_hi[b] = b2 * b2 * (1.0 - 1e-12);
            }
        }
        __syncthreads();
    }

    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int Lm1 = p.Lm1;

    // per-ray state: state i of the current ray.  Inside the stepping loop state i is always a KEPT state
    // (it passed the guard of sidtracemodule.c:36; ray.py:99-102 keeps [:n]).
    long long ray = -1;
    int i = 0, ind = 0;                      // ind: cell of state i (valid while the ray is alive)
    double x = 0, y = 0, kx = 0, ky = 0;
    double st_r = p.st, sst_r = p.sst;      // per-ray step (STRAY)
    double *tp = nullptr;
    // extras
    int turn = -1;
    double kprev = 0, lax = 0, ph = 0, a2 = 1.0;
    int binprev = -1;        // HIST: flux bin of the previous kept state = bin of the current run
    double binacc = 0.0;     // HIST: weight gathered for that bin since the ray entered it
    double r2lo = 1.0, r2hi = 0.0;   // HIST: squared-radius annulus of that bin (empty = unknown)
    // crossing.  The receiver coordinate of a state is c = c0p + (g - g0) * dg with g = y or x (ray.py:105-106).
    // Per step only its offset d = c - val is carried; everything a capture needs beyond the registers of
    // the segment's start state is re-read from global memory at capture time (once per ray), and what a
    // capture produces goes straight to the output arrays (the LOOKUP pass keeps it in registers too).
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const bool axis_y = p.cross_axis != 2;
    const double dg = axis_y ? p.dy : p.dx;
    double c0p = 0, g0 = 0, dcur = 0;
    int cidx = -1, ncross = 0, neq = 0;
    int sgn_or = 0, sgn_and = -1;   // OR / AND of the high words of every kept state's line offset: any below / any above the line
    // state n-1 = the last kept state, assigned once when the ray ends (retire() extrapolates from it)
    double pxg = 0, pyg = 0, pkx = 0, pky = 0, plax = 0, pph = 0;
    int pind = 0;
    // lookup: launch point, the captured crossing (until it is inserted into the polyline), post-insertion
    // running sums, last two polyline points, targets
    double x0p = 0, y0p = 0, xg0 = 0, yg0 = 0;
    double cx = qnan, cy = qnan, clax = qnan, cph = qnan, cs = 0, ckn = 0;
    bool cap_seg = false;    // a crossing point was captured in the segment that ends at state i
    double lax2 = 0, ph2 = 0;
    double sa_ph = 0, sa_x = 0, sa_y = 0, sb_ph = 0, sb_x = 0, sb_y = 0;
    int npts = 0;
    double tA = 0, tB = 0, oxA = qnan, oyA = qnan, oxB = qnan, oyB = qnan;
    bool doneA = false, doneB = false;

    // sidtracemodule.c:36 -- guard on a state with index j:  j < L-1  and  0 <= x < nx-1  and  0 <= y < ny-1
    // (false for NaN).  The O(rays)-output modes are bound by instruction issue, so the four FP64 compares
    // are done on the integer pipe instead, exactly: with c = trunc(v) (saturating, NaN -> 0),
    //   0 <= v < n   <=>   v is +finite (or -0.0)  and  (unsigned)c < n        (n integer < 2^31)
    // "+finite": high word, read as unsigned, below 0x7ff00000 (excludes negatives, inf, NaN);
    // -0.0 (which compares >= 0 in the reference) is admitted explicitly.  Returns the cell (:38).
    auto guard = [&](double xs, double ys, int j, int &cell) -> bool {
        const int ix = __double2int_rz(xs), iy = __double2int_rz(ys);
        cell = iy * p.csy + ix * p.csx;
        return (j < Lm1) & in_range(xs, ix, p.nxm1) & in_range(ys, iy, p.nym1);
    };

    // Take ray r: load its launch state, write state 0, reset the per-ray outputs.  False when the launch
    // state already fails the guard (the ray takes no step: nsteps = 0, end state = launch state).
    auto launch = [&](long long r, long long slot) -> bool {
        ray = r;
        x = p.x0[r]; y = p.y0[r]; kx = p.kx0[r]; ky = p.ky0[r];
        if (STRAY && p.st_ray) { st_r = p.st_ray[r]; sst_r = __dmul_rn(st_r, st_r); }
        i = 0;
        if (TRAJ) {
            if (TILED && p.tile_off) {
                // packed: tile t owns rows tile_off[t] .. tile_off[t+1] (sized for its longest ray); never
                // step beyond them, whatever the plan was made from
                const long long t = slot / p.tile, row0 = p.tile_off[t];
                tp = p.traj + (row0 * p.tile + (slot - t * p.tile)) * 4;
                Lm1 = min(p.L - 1, (int)(p.tile_off[t + 1] - row0) - 1);
            } else {
                tp = TILED ? p.traj + ((slot / p.tile) * (long long)p.L * p.tile + (slot % p.tile)) * 4
                           : p.traj + r * p.stride;
            }
            st_state<true>(tp, x, y, kx, ky);
        }
        if (EXTRAS) { turn = -1; lax = 0; ph = 0; }
        if (CROSS) {
            c0p = axis_y ? p.y0p[r] : p.x0p[r];
            g0 = axis_y ? y : x;
            dcur = c0p - p.cross_val;            // c0p + (g0 - g0) * dg - val
            cidx = -1; ncross = 0; neq = 0; sgn_or = 0; sgn_and = -1;
            if (LOOKUP) {
                x0p = p.x0p[r]; y0p = p.y0p[r]; xg0 = x; yg0 = y;
                cx = cy = clax = cph = qnan;
                tA = p.lookup_ph[r]; tB = p.lookup_ph[p.nrays + r];
                oxA = oyA = oxB = oyB = qnan;
                doneA = doneB = cap_seg = false;
                npts = 0; lax2 = ph2 = 0;
            }
        }
        if (HIST) {
            const double a = p.amp ? p.amp[r] : 1.0;
            a2 = __dmul_rn(a, a);
            binprev = -1; binacc = 0.0; r2lo = 1.0; r2hi = 0.0;
        }
        return guard(x, y, 0, ind);
    };

    // crossing capture (out of line, see cross_capture); `at` is the index Ray.addpoint reports
    auto capture = [&](double sx, double sy, double skx, double sky, int sind, double slax, double sph, double st1, int at) -> bool {
        const CrossCtx c{p.field, p.x0p, p.y0p, p.x0, p.y0, p.cross_state, p.nrays, p.dx, p.dy, p.cross_val, axis_y};
        CrossPoint q;
        if (!cross_capture(c, ray, sx, sy, skx, sky, sind, slax, sph, st1, &q)) return false;
        cidx = at;
        if (LOOKUP) { cx = q.x; cy = q.y; clax = q.lax; cph = q.ph; cs = q.s; ckn = q.kn; }
        return true;
    };

    // append a point (phase, x, y) to the ray's polyline and resolve the targets it brackets:
    // interp1d picks the first node b >= 1 with t <= ph[b] and interpolates on (b-1, b)
    auto interp_seg = [&](double t, double &ox, double &oy) {
        // interp1d._call_linear (SciPy >= 1.12): ((t - lo)/(hi - lo))*y_hi + ((hi - t)/(hi - lo))*y_lo
        const double dph = sb_ph - sa_ph;
        const double wa = (t - sa_ph) / dph, wb = (sb_ph - t) / dph;
        ox = wa * sb_x + wb * sa_x;
        oy = wa * sb_y + wb * sa_y;
    };
    auto add_point = [&](double phq, double xq, double yq) {
        sa_ph = sb_ph; sa_x = sb_x; sa_y = sb_y;
        sb_ph = phq; sb_x = xq; sb_y = yq;
        if (++npts >= 2) {
            if (!doneA && tA <= phq) { interp_seg(tA, oxA, oyA); doneA = true; }
            if (!doneB && tB <= phq) { interp_seg(tB, oxB, oyB); doneB = true; }
        }
    };

    // per-ray results: here i = n, (x, y, kx, ky) = state n -- the first out-of-domain state or the cap
    // (sidtracemodule.c:50) -- and (pxg, ..., pph) = state n-1 when n >= 1.
    // Ray-granular schedules retire a ray the moment it ends; the TILED schedule retires the
    // whole tile at once, so nsteps / end states leave the warp as coalesced 128 B / 256 B rows
    // (which is what makes writing them straight into mapped host memory cheap).
    auto retire = [&]() {
        if (HIST) {
            if (binacc != 0.0) hist_add(sh_hist, binprev, binacc);
            binacc = 0.0;
        }
        if (p.nsteps) p.nsteps[ray] = i;
        if (p.endst) {
            p.endst[0 * p.nrays + ray] = x;
            p.endst[1 * p.nrays + ray] = y;
            p.endst[2 * p.nrays + ray] = kx;
            p.endst[3 * p.nrays + ray] = ky;
        }
        if (EXTRAS) {
            if (p.turn) p.turn[ray] = turn;
            if (p.turn_state && turn < 0) {           // a turning point wrote its state when it was found
                p.turn_state[0 * p.nrays + ray] = qnan; p.turn_state[1 * p.nrays + ray] = qnan;
                p.turn_state[2 * p.nrays + ray] = qnan; p.turn_state[3 * p.nrays + ray] = qnan;
            }
            if (p.lax_end) p.lax_end[ray] = lax;
            if (p.ph_end) p.ph_end[ray] = ph;
        }
        if (CROSS) {
            // states 0..n-1 are kept
            int count = 0, flags = 0;
            if (neq > 0) {                              // ray.py:142-143: exact hit(s), nothing inserted
                count = neq; flags = 1;
            } else if (sgn_or < 0 && sgn_and >= 0) {    // kept states on both sides of the line, ray.py:144-146.  (A zero
                                                        // offset would count for its sign bit's side -- but then neq > 0
                                                        // and the branch above was taken; NaN offsets find no crossing
                                                        // on either path.)
                count = cidx >= 0 ? ncross : 0;
            } else if (i >= 1) {                        // ray.py:147-148: heading for the line past the end
                const double cl = c0p + ((axis_y ? pyg : pxg) - g0) * dg;
                if ((p.cross_val - cl) * (axis_y ? pky : pkx) > 0.0 && capture(pxg, pyg, pkx, pky, pind, plax, pph, 0.0, i)) {
                    count = 1; flags = 2;
                }
            }
            if (LOOKUP) {
                if (flags == 2) add_point(cph, cx, cy);            // point appended past the last kept state
                if (npts >= 2) {                                    // beyond the last node: last segment extrapolates
                    if (!doneA) interp_seg(tA, oxA, oyA);
                    if (!doneB) interp_seg(tB, oxB, oyB);
                }
                if (p.lookup_xy) {
                    p.lookup_xy[0 * p.nrays + ray] = oxA; p.lookup_xy[1 * p.nrays + ray] = oyA;
                    p.lookup_xy[2 * p.nrays + ray] = oxB; p.lookup_xy[3 * p.nrays + ray] = oyB;
                }
            }
            if (count == 0) {
                cidx = -1;
                if (p.cross_state)
                    for (int k = 0; k < 6; ++k) p.cross_state[k * p.nrays + ray] = qnan;
            }
            if (p.cross_count) p.cross_count[ray] = count;
            if (p.cross_index) p.cross_index[ray] = cidx;
            if (p.cross_flags) p.cross_flags[ray] = flags;
        }
    };

    // One step of the reference loop from the KEPT state i: per-state bookkeeping, state i+1, the
    // guard on state i+1.  True when state i+1 is kept as well (the registers then hold it); false
    // when the ray has ended: the registers hold state n = i+1 (written, not kept) and p* state n-1.
    auto step = [&]() -> bool {
        const double st = STRAY ? st_r : p.st, sst = STRAY ? sst_r : p.sst;
        // :38  nearest lower-left cell, no interpolation
        double2 m = make_double2(0.0, 0.0);
        double4 h = make_double4(0.0, 0.0, 0.0, 0.0);
        if (HOIST) h = ld_cell4<TRAJ>(&p.field4[ind]);
        else m = ld_cell2<TRAJ>(&p.field[ind]);

        double kavr = 0, sk = 0, kcur = 0;
        if (EXTRAS || HIST) {
            // close segment (i-1, i)
            kcur = kmod(kx, ky);
            if (i > 0) {
                kavr = __dmul_rn(__dadd_rn(kprev, kcur), 0.5);   // ray.py:118
                sk = __dmul_rn(st, kavr);
                if (EXTRAS) {
                    const double lnew = __dadd_rn(lax, sk);                  // :120 cumsum
                    const double dl = __dsub_rn(lnew, lax);                  // :124 diff
                    ph = __dadd_rn(ph, __dmul_rn(dl, kavr));                 // :125
                    lax = lnew;
                }
                // a ray stays in one flux bin for many consecutive steps: gather the run in a
                // register and touch shared memory only when the bin changes
                if (HIST) binacc = __dadd_rn(binacc, __dmul_rn(a2, sk));
            }
            kprev = kcur;
            if (HIST) {
                // flux label of the segment START (oracle_deposit).  While the squared radius stays
                // strictly inside the current bin's annulus (shrunk by 1e-12, far more than the
                // rounding of the exact formula) the bin cannot have changed: no sqrt, no division.
                const double xp = __dadd_rn(p.xorg, __dmul_rn(x, p.dx));
                const double yp = __dadd_rn(p.yorg, __dmul_rn(y, p.dy));
                const double r2 = __dadd_rn(__dmul_rn(xp, xp), __dmul_rn(yp, yp));
                if (!p.hist_annulus || !(r2 > r2lo && r2 < r2hi)) {
                    // a ray that leaves its bin is almost always in one of the two neighbouring bins: their
                    // certain-annuli come from shared memory (two loads + two compares each) before the exact
                    // label (square root, two exact quotients, conversion) is computed at all
                    int bnew = -1;
                    bool known = false;
                    if (p.hist_table && binprev >= 0) {
                        const int bu = binprev + 1, bd = binprev - 1;
                        if (bu < p.nbins) {
                            const double lo = sh_lo[bu], hi = sh_hi[bu];
                            if (r2 > lo && r2 < hi) { bnew = bu; r2lo = lo; r2hi = hi; known = true; }
                        }
                        if (!known && bd >= 0) {
                            const double lo = sh_lo[bd], hi = sh_hi[bd];
                            if (r2 > lo && r2 < hi) { bnew = bd; r2lo = lo; r2hi = hi; known = true; }
                        }
                    }
                    double rho = 0.0;
                    if (!known) {
                        rho = div_const<DIVC>(__dsqrt_rn(r2), p.r_lcs, p.r_rlcs);
                        if (rho < p.rho_max) {
                            const int b = __double2int_rz(div_const<DIVC>(__dmul_rn(rho, (double)p.nbins), p.rho_max, p.r_rhomax));
                            if (b >= 0 && b < p.nbins) bnew = b;
                        }
                    }
                    if (bnew != binprev) {
                        if (binacc != 0.0) hist_add(sh_hist, binprev, binacc);
                        binprev = bnew;
                        binacc = 0.0;
                    }
                    // annulus of the bin just found, in squared physical radius
                    const double dr = p.hist_dr;
                    if (known) {
                        // bounds already taken from the table
                    } else if (!p.hist_annulus) {
                        // bins narrower than a step: the test would fail most of the time
                    } else if (bnew >= 0) {
                        const double a = dr * (double)bnew, b2 = dr * (double)(bnew + 1);
                        r2lo = a * a * (1.0 + 1e-12);
                        r2hi = b2 * b2 * (1.0 - 1e-12);
                    } else if (rho >= p.rho_max) {
                        const double a = p.rho_max * p.r_lcs;
                        r2lo = a * a * (1.0 + 1e-12);
                        r2hi = __longlong_as_double(0x7ff0000000000000LL);
                    } else {
                        r2lo = 1.0; r2hi = 0.0;       // NaN / odd cases: always take the exact path
                    }
                }
            }
        }

        if (CROSS) {
            // state i is kept; lax, ph are its running sums; dcur = its receiver coordinate - val
            // (the exact-hit test `dcur == 0` of ray.py:142-143 moved behind the product test at the bottom of the
            // step: one DMUL + one DSETP per step cover it together with the sign change; an FP64 compare costs
            // two dispatch cycles, and these passes are dispatch-bound)
            sgn_or |= __double2hiint(dcur);
            sgn_and &= __double2hiint(dcur);
            if (LOOKUP) {
                // the polyline Beam.field_r sees after addpoint: kept states plus the inserted crossing
                // point, phase recomputed (ray.py:205-216, 122-126).  Before the insertion its sums equal
                // lax / ph (same operations); the crossing point splits segment (i-1, i): st[i-1] = s,
                // st[new] = st - s (ray.py:205-210)
                const double xpi = x0p + (x - xg0) * p.dx, ypi = y0p + (y - yg0) * p.dy;
                if (i > 0) {
                    if (cap_seg) {
                        add_point(cph, cx, cy);
                        const double kav2 = (ckn + kcur) / 2;
                        const double l2 = clax + (st - cs) * kav2;
                        ph2 = cph + (l2 - clax) * kav2;
                        lax2 = l2;
                        cap_seg = false;
                    } else {
                        const double l2 = lax2 + sk;
                        ph2 = ph2 + (l2 - lax2) * kavr;
                        lax2 = l2;
                    }
                }
                add_point(ph2, xpi, ypi);
            }
        }

        const double tx = HOIST ? h.x : __dmul_rn(__dmul_rn(sst, m.x), 0.5);   // ((st*st)*Mx)/2
        const double ty = HOIST ? h.y : __dmul_rn(__dmul_rn(sst, m.y), 0.5);
        const double vx = __dadd_rn(tx, __dmul_rn(kx, st));
        const double vy = __dadd_rn(ty, __dmul_rn(ky, st));
        double wx, wy;
        div_pair<DIVC>(vx, vy, p.dx, p.dy, p.rdx, p.rdy, wx, wy);
        const double x1 = __dadd_rn(x, wx);                       // :41
        const double y1 = __dadd_rn(y, wy);                       // :42
        const double kx1 = __dadd_rn(kx, HOIST ? h.z : __dmul_rn(st, m.x));     // :44
        const double ky1 = __dadd_rn(ky, HOIST ? h.w : __dmul_rn(st, m.y));     // :45

        if (EXTRAS) {
            if (!(ky * ky1 > 0.0) && turn < 0 && ((ky > 0.0 && ky1 <= 0.0) || (ky < 0.0 && ky1 >= 0.0))) {
                turn = i + 1;
                if (p.turn_state) {
                    p.turn_state[0 * p.nrays + ray] = x1;  p.turn_state[1 * p.nrays + ray] = y1;
                    p.turn_state[2 * p.nrays + ray] = kx1; p.turn_state[3 * p.nrays + ray] = ky1;
                }
            }
        }
        if (TRAJ) {     // states 0..n are written (sidtracemodule.c:41-45), kept or not
            tp += TILED ? 4 * p.tile : 4;
            st_state<true>(tp, x1, y1, kx1, ky1);
        }
        int ind1;
        const bool ok1 = guard(x1, y1, i + 1, ind1);              // :36 on state i+1
        if (CROSS) {
            const double d1 = (c0p + ((axis_y ? y1 : x1) - g0) * dg) - p.cross_val;
            if (!(dcur * d1 > 0.0)) {                      // rare: a zero offset, a sign change, an underflowed or NaN product
                if (dcur == 0.0) {                         // ray.py:142-143: the kept state i lies on the line
                    if (neq == 0) {
                        cidx = i;
                        const CrossCtx c{p.field, p.x0p, p.y0p, p.x0, p.y0, p.cross_state, p.nrays, p.dx, p.dy, p.cross_val, axis_y};
                        cross_exact_hit(c, ray, x, y, kx, ky, lax, ph);
                    }
                    ++neq;
                }
                // a sign change between two KEPT states (ray.py:144-146); the first one is solved for
                if (ok1 && dcur * d1 < 0.0) {
                    if (ncross == 0 && neq == 0) {
                        const bool got = capture(x, y, kx, ky, ind, lax, ph, st, i + 1);
                        if (LOOKUP) cap_seg = got;
                    }
                    ++ncross;
                }
            }
            dcur = d1;
        }
        if (!ok1) {
            if (CROSS) { pxg = x; pyg = y; pkx = kx; pky = ky; pind = ind; plax = lax; pph = ph; }
            x = x1; y = y1; kx = kx1; ky = ky1;
            ++i;
            return false;
        }
        // (An L1 prefetch of the cell 1-3 steps further along the last cell-to-cell displacement was measured
        // for the O(rays)-output modes, whose top stall is the wait for the gather: 13-18 % SLOWER,
        // profiles/r02/prefetch_dropped.txt.)
        x = x1; y = y1; kx = kx1; ky = ky1; ind = ind1;
        ++i;
        return true;
    };

    // LEAN variants (step counts / end states only: nothing is written per step): the common case of a
    // step -- both dividends in div_pair's guarded exponent range, the new state +finite and inside the
    // grid, below the index limit -- runs in a loop of its own.  It leaves the loop in two ways:
    //   0  a dividend outside the guarded range, seen BEFORE anything is committed: registers still hold the
    //      kept state i, the general step() takes it from there;
    //   the new state failed the quick guard (one unsigned compare of the high word per coordinate, which
    //   also turns away -0.0, next to the cell-range compare) AFTER it was committed: registers hold state
    //   i (already counted); the exact guard decides between
    //   1  the ray has ended (registers = state n, sidtracemodule.c:50) and
    //   2  the state is kept after all (-0.0, or only the caller's index limit was reached).
    // Same operations on the same operands as step(): the results cannot differ.  One ray-step is
    // 39 instead of 43 SASS instructions (+3 % in endpoint mode on the static grid; the persistent schedules keep
    // the plain step() loop).
    constexpr bool LEAN = !TRAJ && !EXTRAS && !HIST && !CROSS && !STRAY && !REFILL;
    auto fast_run = [&](int ilim) -> int {
        if (!LEAN) return 0;
#pragma unroll 1
        for (;;) {
            double tx, ty, sx, sy;
            if (HOIST) {
                const double4 h = ld_cell4<false>(&p.field4[ind]);
                tx = h.x; ty = h.y; sx = h.z; sy = h.w;
            } else {
                const double2 m = ld_cell2<false>(&p.field[ind]);
                tx = __dmul_rn(__dmul_rn(p.sst, m.x), 0.5); ty = __dmul_rn(__dmul_rn(p.sst, m.y), 0.5);
                sx = __dmul_rn(p.st, m.x); sy = __dmul_rn(p.st, m.y);
            }
            const double vx = __dadd_rn(tx, __dmul_rn(kx, p.st));
            const double vy = __dadd_rn(ty, __dmul_rn(ky, p.st));
            double wx, wy;
            if (DIVC) {
                const unsigned ex = ((unsigned)__double2hiint(vx) & 0x7ff00000u) - (623u << 20);
                const unsigned ey = ((unsigned)__double2hiint(vy) & 0x7ff00000u) - (623u << 20);
                if (max(ex, ey) > (800u << 20)) return 0;
                const double qx = __dmul_rn(vx, p.rdx), qy = __dmul_rn(vy, p.rdy);
                const double rx = __fma_rn(-qx, p.dx, vx), ry = __fma_rn(-qy, p.dy, vy);
                wx = __fma_rn(rx, p.rdx, qx);
                wy = __fma_rn(ry, p.rdy, qy);
            } else {
                wx = __ddiv_rn(vx, p.dx);
                wy = __ddiv_rn(vy, p.dy);
            }
            x = __dadd_rn(x, wx); y = __dadd_rn(y, wy);
            kx = __dadd_rn(kx, sx); ky = __dadd_rn(ky, sy);
            ++i;
            const int ix = __double2int_rz(x), iy = __double2int_rz(y);
            ind = iy * p.csy + ix * p.csx;
            const bool okx = ((unsigned)__double2hiint(x) < 0x7ff00000u) & ((unsigned)ix < (unsigned)p.nxm1);
            const bool oky = ((unsigned)__double2hiint(y) < 0x7ff00000u) & ((unsigned)iy < (unsigned)p.nym1);
            if (!(okx & oky & (i < ilim))) break;
        }
        return guard(x, y, i, ind) ? 2 : 1;
    };

    if (TILED || !REFILL) {
        // ---- whole-ray schedules: claim, trace to the end in a plain per-thread loop, retire.
        // No vote per step: lanes that finish early wait at the reconvergence point after the
        // inner loop, which for a tile of neighbouring rays costs a few steps.
        long long r = REFILL ? 0 : (long long)blockIdx.x * BLOCK + threadIdx.x;
        const bool ordered = p.perm != nullptr;
        for (;;) {
            if (REFILL) {   // TILED: one aggregated claim of 32 consecutive rays (slots) per warp
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(p.counter, 32ull);
                base = __shfl_sync(FULL, base, 0);
                if ((long long)base >= (ordered ? p.nslots : p.nrays)) break;
                r = (long long)base + lane;
            }
            const long long slot = r;
            if (ordered) r = slot < p.nslots ? (long long)p.perm[slot] : -1;
            const bool has = r >= 0 && r < p.nrays;
            if (has && launch(r, slot)) {
                if (LEAN) {
                    for (;;) {
                        const int why = fast_run(Lm1);
                        if (why == 1 || (why == 0 && !step())) break;
                    }
                } else {
                    // the fused-sum and histogram passes gain 4 % / 1-5 % from a loop unrolled by two (ptxas then
                    // rotates the state registers instead of copying them); trajectories, crossing and lookup lose
                    // 3-4 % (profiles/r02/crossopt_unr2.txt)
                    if ((EXTRAS || HIST) && !CROSS && !TRAJ) {
#pragma unroll 2
                        while (step()) { }
                    } else {
                        while (step()) { }
                    }
                }
            }
            if (TILED) __syncwarp();
            if (has) retire();
            if (!REFILL) break;
        }
    } else {
        // ---- ray-granular refill: a lane whose ray ended claims the next one.  The warp votes
        // once per kBatch steps; when at least kRefillMin lanes are idle (or all of them) they are
        // refilled with ONE aggregated atomic -- claiming rays 8+ at a time keeps the lanes of a
        // warp on neighbouring rays (gather locality) at the price of <= 7 idle lanes.  Measured
        // (profiles/r01/kernel_sweep_refill_batch_threshold.txt): 16 steps / 8 lanes beats
        // 8 steps / 1 lane by 3 % (clean beam), 9 % (trapped rays), 19 % (cutoff sweep).
#ifndef STB_BATCH
#define STB_BATCH 16
#endif
#ifndef STB_REFILL_MIN
#define STB_REFILL_MIN 8
#endif
        constexpr int kBatch = STB_BATCH;
        constexpr int kRefillMin = STB_REFILL_MIN;   // idle lanes that trigger a refill (or the whole warp idle)
        bool active = false, more = true;
        for (;;) {
            unsigned act = __ballot_sync(FULL, active);
            if (more && (act == 0 || __popc(~act) >= kRefillMin)) {
                const unsigned idle = ~act;
                const int leader = __ffs(idle) - 1;
                const int want = __popc(idle);
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(p.counter, (unsigned long long)want);
                base = __shfl_sync(FULL, base, leader);
                const long long nclaim = p.perm ? p.nslots : p.nrays;
                if (!active) {
                    const long long slot = (long long)base + __popc(idle & ((1u << lane) - 1u));
                    const long long r = slot >= nclaim ? -1 : p.perm ? (long long)p.perm[slot] : slot;
                    if (r >= 0 && r < p.nrays) {
                        if (launch(r, slot)) active = true;
                        else retire();                      // a ray that takes no step
                    }
                }
                more = (long long)(base + want) < nclaim;
                act = __ballot_sync(FULL, active);
            }
            if (act == 0) {
                if (!more) break;
                continue;                                   // every claimed ray ended at launch: claim again
            }
            if (active) {
                // (the LEAN loop with a per-batch index limit was measured here too: 10 % slower than plain
                // step() calls, profiles/r02/lean_sweep_config3.txt)
#pragma unroll 1
                for (int s = 0; s < kBatch; ++s) {
                    if (!step()) { retire(); active = false; break; }
                }
            }
        }
    }

    if (HIST) {
        __syncthreads();
        for (int b = threadIdx.x; b < p.nbins; b += BLOCK) {
            const double v = sh_hist[b];
            if (v != 0.0) atomicAdd(&p.hist[b], v);
        }
    }
}

// x-major cells -> y-major copy (out[ix*nym1 + iy] = in[iy*nxm1 + ix]); 32x32 tiles through shared memory
template <typename T>
__global__ void transpose_cells_kernel(const T *__restrict__ in, T *__restrict__ out, int nxm1, int nym1)
{
    __shared__ T tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int ix = bx + threadIdx.x, iy = by + j;
        if (ix < nxm1 && iy < nym1) tile[j][threadIdx.x] = in[(long long)iy * nxm1 + ix];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int ix = bx + j, iy = by + threadIdx.x;
        if (ix < nxm1 && iy < nym1) out[(long long)ix * nym1 + iy] = tile[threadIdx.x][j];
    }
}

// Endpoint mode, static grid, TWO rays per thread.  The kernels without trajectory output are bound by the latency
// of the step's dependent chain (see above): throughput = independent chains in flight / chain latency, and one
// ray per thread caps the chains at the 48 warps/SM the register file allows.  Here a thread carries rays
// r and r + BLOCK of its CTA's 2*BLOCK consecutive rays, so every warp has two independent chains to issue from
// (launch-state loads and per-ray results stay coalesced).  While both rays are alive and in the common case --
// all four dividends in div_pair's guarded exponent range, both new states +finite, inside the grid, below the
// cap -- one loop iteration steps both; anything else is settled by the exact single-ray step / guard, and once
// one ray has ended the other finishes alone.  Same operations on the same operands as step(): bit-identical.
#ifndef STB_MINB_DUAL
#define STB_MINB_DUAL 9       // 56 registers, 36 warps = 72 rays in flight per SM (6 / 8 / 9 / 10 CTAs per SM measured:
#endif                        // 1.418 / 1.285 / 1.255 / 1.254 ms per 10^6 rays; one ray per thread: 1.299 ms)
struct DualRay { double x, y, kx, ky; int ind, i; };

template <bool DIVC, bool HOIST>
__global__ void __launch_bounds__(kBlock, STB_MINB_DUAL) endpoint_dual_kernel(const KParams p)
{
    const long long sA = (long long)blockIdx.x * (2 * kBlock) + threadIdx.x, sB = sA + kBlock;
    const long long nslot = p.perm ? p.nslots : p.nrays;
    const long long rA = sA >= nslot ? -1 : p.perm ? (long long)p.perm[sA] : sA;      // perm: the order the rays are traced in
    const long long rB = sB >= nslot ? -1 : p.perm ? (long long)p.perm[sB] : sB;
    const int Lm1 = p.Lm1;

    auto guard = [&](const DualRay &s) -> bool {       // sidtracemodule.c:36 on state s.i, exact (see trace_kernel)
        const int ix = __double2int_rz(s.x), iy = __double2int_rz(s.y);
        return (s.i < Lm1) & in_range(s.x, ix, p.nxm1) & in_range(s.y, iy, p.nym1);
    };
    auto cell_of = [&](const DualRay &s) -> int { return __double2int_rz(s.y) * p.csy + __double2int_rz(s.x) * p.csx; };
    auto load = [&](long long r, DualRay &s) -> bool {
        if (r < 0 || r >= p.nrays) return false;
        s.x = p.x0[r]; s.y = p.y0[r]; s.kx = p.kx0[r]; s.ky = p.ky0[r]; s.i = 0;
        if (!guard(s)) return false;
        s.ind = cell_of(s);
        return true;
    };
    auto cell = [&](int ind, double &tx, double &ty, double &sx, double &sy) {
#ifdef STB_DIAG_FIXED_CELL     // timing diagnostic only (wrong results): every gather hits the same L1 line
        ind &= 3;
#endif
        if (HOIST) {
            const double4 h = ld_cell4<false>(&p.field4[ind]);
            tx = h.x; ty = h.y; sx = h.z; sy = h.w;
        } else {
            const double2 m = ld_cell2<false>(&p.field[ind]);
            tx = __dmul_rn(__dmul_rn(p.sst, m.x), 0.5); ty = __dmul_rn(__dmul_rn(p.sst, m.y), 0.5);
            sx = __dmul_rn(p.st, m.x); sy = __dmul_rn(p.st, m.y);
        }
    };
    // the exact step from the kept state s.i (the reference's loop body + guard): true when state i+1 is kept too
    auto exact_step = [&](DualRay &s) -> bool {
        double tx, ty, sx, sy;
        cell(s.ind, tx, ty, sx, sy);
        const double vx = __dadd_rn(tx, __dmul_rn(s.kx, p.st)), vy = __dadd_rn(ty, __dmul_rn(s.ky, p.st));
        double wx, wy;
        div_pair<DIVC>(vx, vy, p.dx, p.dy, p.rdx, p.rdy, wx, wy);
        s.x = __dadd_rn(s.x, wx); s.y = __dadd_rn(s.y, wy);
        s.kx = __dadd_rn(s.kx, sx); s.ky = __dadd_rn(s.ky, sy);
        ++s.i;
        if (!guard(s)) return false;
        s.ind = cell_of(s);
        return true;
    };
    auto exp_bits = [](double v) -> unsigned { return ((unsigned)__double2hiint(v) & 0x7ff00000u) - (623u << 20); };
    auto quick_ok = [&](const DualRay &s, int ix, int iy) -> bool {
        return ((unsigned)__double2hiint(s.x) < 0x7ff00000u) & ((unsigned)ix < (unsigned)p.nxm1)
             & ((unsigned)__double2hiint(s.y) < 0x7ff00000u) & ((unsigned)iy < (unsigned)p.nym1) & (s.i < Lm1);
    };
    auto retire = [&](long long r, const DualRay &s) {
        if (r < 0 || r >= p.nrays) return;
        if (p.nsteps) p.nsteps[r] = s.i;
        if (p.endst) {
            p.endst[0 * p.nrays + r] = s.x; p.endst[1 * p.nrays + r] = s.y;
            p.endst[2 * p.nrays + r] = s.kx; p.endst[3 * p.nrays + r] = s.ky;
        }
    };

    DualRay a{}, b{};
    bool la = load(rA, a), lb = load(rB, b);

    // ---- both rays alive
#pragma unroll 1
    while (la & lb) {
        double txa, tya, sxa, sya, txb, tyb, sxb, syb;
        cell(a.ind, txa, tya, sxa, sya);
        cell(b.ind, txb, tyb, sxb, syb);
        const double vxa = __dadd_rn(txa, __dmul_rn(a.kx, p.st)), vya = __dadd_rn(tya, __dmul_rn(a.ky, p.st));
        const double vxb = __dadd_rn(txb, __dmul_rn(b.kx, p.st)), vyb = __dadd_rn(tyb, __dmul_rn(b.ky, p.st));
        if (DIVC && max(max(exp_bits(vxa), exp_bits(vya)), max(exp_bits(vxb), exp_bits(vyb))) > (800u << 20)) {
            la = exact_step(a);                       // nothing committed yet: redo both exactly
            lb = exact_step(b);
            continue;
        }
        double wxa, wya, wxb, wyb;
        if (DIVC) {
            const double qxa = __dmul_rn(vxa, p.rdx), qya = __dmul_rn(vya, p.rdy), qxb = __dmul_rn(vxb, p.rdx), qyb = __dmul_rn(vyb, p.rdy);
            const double exa = __fma_rn(-qxa, p.dx, vxa), eya = __fma_rn(-qya, p.dy, vya);
            const double exb = __fma_rn(-qxb, p.dx, vxb), eyb = __fma_rn(-qyb, p.dy, vyb);
            wxa = __fma_rn(exa, p.rdx, qxa); wya = __fma_rn(eya, p.rdy, qya);
            wxb = __fma_rn(exb, p.rdx, qxb); wyb = __fma_rn(eyb, p.rdy, qyb);
        } else {
            wxa = __ddiv_rn(vxa, p.dx); wya = __ddiv_rn(vya, p.dy); wxb = __ddiv_rn(vxb, p.dx); wyb = __ddiv_rn(vyb, p.dy);
        }
        a.x = __dadd_rn(a.x, wxa); a.y = __dadd_rn(a.y, wya); a.kx = __dadd_rn(a.kx, sxa); a.ky = __dadd_rn(a.ky, sya); ++a.i;
        b.x = __dadd_rn(b.x, wxb); b.y = __dadd_rn(b.y, wyb); b.kx = __dadd_rn(b.kx, sxb); b.ky = __dadd_rn(b.ky, syb); ++b.i;
        const int ixa = __double2int_rz(a.x), iya = __double2int_rz(a.y), ixb = __double2int_rz(b.x), iyb = __double2int_rz(b.y);
        a.ind = iya * p.csy + ixa * p.csx;
        b.ind = iyb * p.csy + ixb * p.csx;
        const bool oka = quick_ok(a, ixa, iya), okb = quick_ok(b, ixb, iyb);
        if (oka & okb) continue;
        la = oka || guard(a);                         // -0.0 passes the reference's `>= 0`
        lb = okb || guard(b);
    }
    // ---- one ray left
    if (la | lb) {
        DualRay &s = la ? a : b;
        bool alive = true;
#pragma unroll 1
        while (alive) {
            double tx, ty, sx, sy;
            cell(s.ind, tx, ty, sx, sy);
            const double vx = __dadd_rn(tx, __dmul_rn(s.kx, p.st)), vy = __dadd_rn(ty, __dmul_rn(s.ky, p.st));
            if (DIVC && max(exp_bits(vx), exp_bits(vy)) > (800u << 20)) { alive = exact_step(s); continue; }
            double wx, wy;
            if (DIVC) {
                const double qx = __dmul_rn(vx, p.rdx), qy = __dmul_rn(vy, p.rdy);
                const double ex = __fma_rn(-qx, p.dx, vx), ey = __fma_rn(-qy, p.dy, vy);
                wx = __fma_rn(ex, p.rdx, qx); wy = __fma_rn(ey, p.rdy, qy);
            } else {
                wx = __ddiv_rn(vx, p.dx); wy = __ddiv_rn(vy, p.dy);
            }
            s.x = __dadd_rn(s.x, wx); s.y = __dadd_rn(s.y, wy); s.kx = __dadd_rn(s.kx, sx); s.ky = __dadd_rn(s.ky, sy); ++s.i;
            const int ix = __double2int_rz(s.x), iy = __double2int_rz(s.y);
            s.ind = iy * p.csy + ix * p.csx;
            alive = quick_ok(s, ix, iy) || guard(s);
        }
    }
    retire(rA, a);
    retire(rB, b);
}

// {ax, ay} -> interleaved double2 per cell.  (A 2x4-cell blocked layout was measured and dropped:
// it did not beat row-major in any mode, profiles/r01/occupancy_layout_sweep.txt.)
__global__ void interleave_kernel(const double *__restrict__ ax, const double *__restrict__ ay,
                                  double2 *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = make_double2(ax[i], ay[i]);
}

// Hoisted table of a (field, st) pair: per cell the four products of sidtracemodule.c:41-45 that do
// not depend on the ray -- the same operations on the same operands the un-hoisted step() performs.
__global__ void hoist_kernel(const double2 *__restrict__ cells, double st, double sst, double4 *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        const double2 m = cells[i];
        double4 h;
        h.x = __dmul_rn(__dmul_rn(sst, m.x), 0.5);
        h.y = __dmul_rn(__dmul_rn(sst, m.y), 0.5);
        h.z = __dmul_rn(st, m.x);
        h.w = __dmul_rn(st, m.y);
        st_state(reinterpret_cast<double *>(out + i), h.x, h.y, h.z, h.w);
    }
}

// Trajectory export: one warp per ray copies the KEPT states 0..n-1 (ray.py:99-102) of a
// trajectory arena -- ray-major (tile == 0) or tiled -- into a packed [total][4] buffer at
// offsets[ray]; 32 B per lane per iteration, coalesced on the packed side.
__global__ void pack_traj_kernel(const double *__restrict__ traj, int tile, long long ray_stride, long long L,
                                 long long nrays, const int *__restrict__ nsteps,
                                 const long long *__restrict__ offsets, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < nrays; r += nwarps) {
        const int n = nsteps[r];
        const double *src = tile ? traj + ((r / tile) * L * tile + (r % tile)) * 4 : traj + r * ray_stride;
        const long long step = tile ? 4LL * tile : 4;
        double *dst = out + offsets[r] * 4;
        for (int j = lane; j < n; j += 32) {
            const double4 v = *reinterpret_cast<const double4 *>(src + j * step);
            st_state(dst + 4LL * j, v.x, v.y, v.z, v.w);
        }
    }
}

// The same from a packed tiled arena (stb_traj_plan).
__global__ void pack_plan_kernel(const double *__restrict__ traj, int tile, const int *__restrict__ slot_of,
                                 const long long *__restrict__ tile_off, long long nrays, const int *__restrict__ nsteps,
                                 const long long *__restrict__ offsets, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < nrays; r += nwarps) {
        const int n = nsteps[r];
        const long long slot = slot_of[r], t = slot / tile;
        const double *src = traj + (tile_off[t] * tile + (slot - t * tile)) * 4;
        const long long step = 4LL * tile;
        double *dst = out + offsets[r] * 4;
        for (int j = lane; j < n; j += 32) {
            const double4 v = *reinterpret_cast<const double4 *>(src + j * step);
            st_state(dst + 4LL * j, v.x, v.y, v.z, v.w);
        }
    }
}

// FP64 peak: 8 independent DFMA chains per thread
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            c0 = __fma_rn(c0, a, b); c1 = __fma_rn(c1, a, b); c2 = __fma_rn(c2, a, b); c3 = __fma_rn(c3, a, b);
            c4 = __fma_rn(c4, a, b); c5 = __fma_rn(c5, a, b); c6 = __fma_rn(c6, a, b); c7 = __fma_rn(c7, a, b);
        }
    }
    const double s = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// streaming 32-byte stores, fully coalesced
__global__ void store_bw_kernel(double *out, long long nvec)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < nvec; i += step) st_state(out + 4 * i, 1.0, 2.0, 3.0, (double)i);
}

// The store pattern of the static tiled trajectory kernel without its arithmetic: CTA b owns tile b,
// thread t writes `steps` states at stride 4*T doubles (one 8 KB row per CTA-step with T = 256), with
// `chain` dependent DFMAs between two stores (0 = stores back to back).  Tells apart "the pattern
// bounds the trajectory kernel" from "its dependency chain does".
__global__ void store_pattern_kernel(double *out, long long rows_per_tile, int steps, int chain, double a, double b)
{
    const int T = blockDim.x;
    double *tp = out + ((long long)blockIdx.x * rows_per_tile * T + threadIdx.x) * 4;
    double v = threadIdx.x;
    for (int i = 0; i < steps; ++i) {
        for (int c = 0; c < chain; ++c) v = __fma_rn(v, a, b);
        st_state(tp, v, 2.0, 3.0, (double)i);
        tp += 4LL * T;
    }
}

}  // namespace stb
