/*
 * sidtrace_oracle.c -- CPU restatement of the reference's per-ray integrator.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rayrtrace_b200/)
 * may import, link or execute this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the known-answer vectors G1..G6 of SURVEY.md section 4.1 (generated from
 *       the compiled, unmodified reference), committed under tests/golden/,
 *   (2) the unmodified reference C extension compiled into oracle/_ref/ (when
 *       present), bit for bit on seeded random fields,
 *   (3) fixtures produced by the reference's own ray.py / den.py
 *       (tests/golden/make_golden.py).
 *
 * Build: plain x86-64, -O2, -ffp-contract=off, no -march=native, so that no
 * FMA contraction can occur (the reference extension is built the same way by
 * setuptools; its gcc -O2 disassembly has no FMA).
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <stdlib.h>

/* ------------------------------------------------------------------------ */
/* sidtracemodule.c:36-50 -- the hot loop.                                   */
/* x, y, kx, ky: length L, element 0 = launch state, mutated in place.       */
/* Returns i = number of steps taken = index of last written state.          */
/* Evaluation order follows C's left-to-right parse of                       */
/*   xx[i+1] = xx[i] + (st * st * Mx[ind] / 2 + kxx[i] * st) / dx;           */
/* i.e. ((st*st)*M)/2, then + (k*st), then /dx, then x + (...).              */
/* ------------------------------------------------------------------------ */
int64_t oracle_trace(double *x, double *y, double *kx, double *ky, int64_t L,
                     double st, double dx, double dy, int nx, int ny,
                     const double *ax, const double *ay)
{
    int64_t i = 0;
    const int nxm1 = nx - 1;
    const int nym1 = ny - 1;
    while (i < L - 1 && x[i] >= 0 && x[i] < nxm1 && y[i] >= 0 && y[i] < nym1) {
        int ind = (int)y[i] * nxm1 + (int)x[i];          /* :38 */
        double mx = ax[ind];
        double my = ay[ind];
        x[i + 1] = x[i] + (st * st * mx / 2 + kx[i] * st) / dx;   /* :41 */
        y[i + 1] = y[i] + (st * st * my / 2 + ky[i] * st) / dy;   /* :42 */
        kx[i + 1] = kx[i] + st * mx;                             /* :44 */
        ky[i + 1] = ky[i] + st * my;                             /* :45 */
        i++;                                                      /* :47 */
    }
    return i;                                                     /* :50 */
}

/* sidtracemodule.c:53-69 -- layout probe: a[i*nx + j].  NOTE the reference
 * multiplies by nx, not nx-1; reproduced as is. */
double oracle_check(int nx, int ny, int i, int j, const double *a)
{
    (void)ny;
    return a[i * nx + j];
}

/* ------------------------------------------------------------------------ */
/* Batch driver: the loop of beam.py:340-342 (one Ray per start point) with   */
/* ray.py:19-25 buffer set-up hoisted out.  Output layout matches the CUDA    */
/* product so that parity tests can memcmp:                                   */
/*   traj  : [ray][step][4] doubles (x, y, kx, ky), ray stride = L*4          */
/*           (may be NULL -> a private scratch of 4*L doubles per thread)      */
/*   nsteps: [ray] int32                                                      */
/*   endst : [4][nrays] doubles = state at index i (the first out-of-domain    */
/*           state, or index L-1 at the cap)    (may be NULL)                  */
/* st_ray may be NULL (use scalar st) or a per-ray step (frequency sweeps).    */
/* ------------------------------------------------------------------------ */
typedef struct {
    int64_t nrays, L;
    const double *x0, *y0, *kx0, *ky0;
    double st; const double *st_ray; double dx, dy;
    int nx, ny; const double *ax, *ay;
    double *traj; int32_t *nsteps; double *endst;
    int64_t next;            /* shared work counter (chunks of 64 rays) */
    int rc;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *jb = (batch_job *)arg;
    const int64_t L = jb->L, nrays = jb->nrays;
    /* per-thread SoA working buffers, as ray.py:19-22 allocates per ray */
    double *buf = (double *)malloc((size_t)L * 4 * sizeof(double));
    if (!buf) { __atomic_store_n(&jb->rc, -2, __ATOMIC_RELAXED); return NULL; }
    double *bx = buf, *by = buf + L, *bkx = buf + 2 * L, *bky = buf + 3 * L;
    for (;;) {
        int64_t r0 = __atomic_fetch_add(&jb->next, 64, __ATOMIC_RELAXED);
        if (r0 >= nrays) break;
        int64_t r1 = r0 + 64 < nrays ? r0 + 64 : nrays;
        for (int64_t r = r0; r < r1; ++r) {
            bx[0] = jb->x0[r]; by[0] = jb->y0[r]; bkx[0] = jb->kx0[r]; bky[0] = jb->ky0[r];
            double s = jb->st_ray ? jb->st_ray[r] : jb->st;
            int64_t n = oracle_trace(bx, by, bkx, bky, L, s, jb->dx, jb->dy,
                                     jb->nx, jb->ny, jb->ax, jb->ay);
            if (jb->nsteps) jb->nsteps[r] = (int32_t)n;
            if (jb->endst) {
                jb->endst[0 * nrays + r] = bx[n];
                jb->endst[1 * nrays + r] = by[n];
                jb->endst[2 * nrays + r] = bkx[n];
                jb->endst[3 * nrays + r] = bky[n];
            }
            if (jb->traj) {
                double *t = jb->traj + (size_t)r * (size_t)L * 4;
                for (int64_t j = 0; j <= n; ++j) {
                    t[4 * j + 0] = bx[j];
                    t[4 * j + 1] = by[j];
                    t[4 * j + 2] = bkx[j];
                    t[4 * j + 3] = bky[j];
                }
            }
        }
    }
    free(buf);
    return NULL;
}

int oracle_trace_batch(int64_t nrays, int64_t L,
                       const double *x0, const double *y0,
                       const double *kx0, const double *ky0,
                       double st, const double *st_ray, double dx, double dy,
                       int nx, int ny, const double *ax, const double *ay,
                       double *traj, int32_t *nsteps, double *endst,
                       int nthreads)
{
    if (L < 1) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    batch_job jb = { nrays, L, x0, y0, kx0, ky0, st, st_ray, dx, dy, nx, ny,
                     ax, ay, traj, nsteps, endst, 0, 0 };
    if (nthreads == 1) {
        batch_worker(&jb);
        return jb.rc;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return -2;
    int started = 0;
    for (int t = 0; t < nthreads; ++t) {
        if (pthread_create(&th[t], NULL, batch_worker, &jb) != 0) break;
        ++started;
    }
    if (started == 0) batch_worker(&jb);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return jb.rc;
}

/* ------------------------------------------------------------------------ */
/* ray.py:113-126 -- |k|, path length `lax` and eikonal phase `ph` of the     */
/* KEPT part of a trajectory (states 0..n-1; the caller drops state n,        */
/* ray.py:99-102).  st[j] = st for j < n-1, st[n-1] = 0 (ray.py:110-111).     */
/*   k[j]   = sqrt(kx^2 + ky^2)                          (:115)               */
/*   kavr_j = (k[j] + k[j+1]) / 2                        (:118, :123)         */
/*   lax    = cumsum([0, st*kavr_0, st*kavr_1, ...])     (:120)               */
/*   dl     = diff(lax)                                  (:124)               */
/*   ph     = cumsum([0, dl_0*kavr_0, ...])              (:125)               */
/* np.cumsum is a sequential left-to-right sum; np.diff(lax)[j] is            */
/* lax[j+1]-lax[j], which is NOT bitwise st*kavr_j in general, so it is        */
/* recomputed from the running sum exactly as numpy does.                     */
/* traj: [n][4] interleaved.  lax, ph: length n (may be NULL).                */
/* Returns ph[n-1] (0 if n < 1).                                              */
/* ------------------------------------------------------------------------ */
double oracle_lax_ph(const double *traj, int64_t n, double st,
                     double *lax, double *ph)
{
    if (n < 1) return 0.0;
    double l = 0.0, p = 0.0;
    if (lax) lax[0] = 0.0;
    if (ph) ph[0] = 0.0;
    double kxp = traj[2], kyp = traj[3];
    double kprev = sqrt(kxp * kxp + kyp * kyp);
    for (int64_t j = 0; j + 1 < n; ++j) {
        double kxn = traj[4 * (j + 1) + 2], kyn = traj[4 * (j + 1) + 3];
        double knext = sqrt(kxn * kxn + kyn * kyn);
        double kavr = (kprev + knext) / 2;
        double lnew = l + st * kavr;
        double dl = lnew - l;
        p = p + dl * kavr;
        l = lnew;
        if (lax) lax[j + 1] = l;
        if (ph) ph[j + 1] = p;
        kprev = knext;
    }
    return p;
}

/* ------------------------------------------------------------------------ */
/* Turning point (defined as a pure function of reference        */
/* trajectories; the reference has no such function): first index j in          */
/* [1, n] at which ky[j] has the opposite strict sign of ky[j-1] or is zero      */
/* while ky[j-1] is not.  Returns -1 if none.                                   */
/* ------------------------------------------------------------------------ */
int64_t oracle_turning_point(const double *traj, int64_t n)
{
    for (int64_t j = 1; j <= n; ++j) {
        double a = traj[4 * (j - 1) + 3], b = traj[4 * j + 3];
        if ((a > 0 && b <= 0) || (a < 0 && b >= 0)) return j;
    }
    return -1;
}

/* ------------------------------------------------------------------------ */
/* Deposition-profile EXTENSION oracle (no reference code).      */
/* For every kept segment j -> j+1 (j+1 <= n-1) of a reference trajectory:     */
/*   w   = amp^2 * dl,  dl = st * (k[j]+k[j+1])/2           (ray.py:117-120)   */
/*   rho = sqrt(xp^2 + yp^2) / r_lcs at the segment START, in physical cm:     */
/*         xp = xorg + xg*dx, yp = yorg + yg*dy (equidistant grid, ray.py:70)  */
/*   bin = (int)(rho * nbins / rho_max); dropped if rho >= rho_max             */
/* hist: nbins doubles, accumulated (not cleared).                             */
/* ------------------------------------------------------------------------ */
void oracle_deposit(const double *traj, int64_t n, double st, double amp,
                    double xorg, double yorg, double dx, double dy,
                    double r_lcs, double rho_max, int nbins, double *hist)
{
    if (n < 2) return;
    double kprev = sqrt(traj[2] * traj[2] + traj[3] * traj[3]);
    for (int64_t j = 0; j + 1 < n; ++j) {
        double kxn = traj[4 * (j + 1) + 2], kyn = traj[4 * (j + 1) + 3];
        double knext = sqrt(kxn * kxn + kyn * kyn);
        double kavr = (kprev + knext) / 2;
        double dl = st * kavr;
        double xp = xorg + traj[4 * j + 0] * dx;
        double yp = yorg + traj[4 * j + 1] * dy;
        double rho = sqrt(xp * xp + yp * yp) / r_lcs;
        if (rho < rho_max) {
            int b = (int)(rho * nbins / rho_max);
            if (b >= 0 && b < nbins) hist[b] += amp * amp * dl;
        }
        kprev = knext;
    }
}

