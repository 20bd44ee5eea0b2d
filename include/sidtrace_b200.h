/*
 * sidtrace_b200.h -- C ABI of the B200-native replacement for the reference's
 * `sidtrace` extension (AVSidorov/RAYRTRACE, sidtracemodule.c).
 *
 * Plain C, plain pointers and sizes; no Python, NumPy or torch types.  This is
 * the surface a maintainer binds from the reference side (ctypes / cffi / a
 * CPython shim -- see INTEGRATION.md).  Every entry point returns an integer
 * status (STB_OK == 0) and never throws across the ABI; the message of the last
 * failure on the calling thread is available from stb_last_error().
 *
 * The library is CUDA-only (sm_100a).  There is no CPU fallback: every compute
 * entry point fails with STB_ERR_CUDA when no usable device is present.
 *
 * Reference interface replaced (file:line in /root/reference):
 *   sidtrace.trace(x,y,kx,ky,st,dx,dy,nx,ny,ax,ay) -> int   sidtracemodule.c:8-51
 *       per-ray call site                                    ray.py:80-83
 *       per-beam call site (one Ray per start point)         beam.py:340-342
 *   sidtrace.check(nx,ny,i,j,a) -> float                     sidtracemodule.c:53-69
 *   method table / module init                               sidtracemodule.c:71-89
 *   build                                                    setup.py:5-12
 */
#ifndef SIDTRACE_B200_H
#define SIDTRACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STB_ABI_VERSION 3

/* status codes */
#define STB_OK            0
#define STB_ERR_ARG      -1   /* bad argument (NULL, bad shape, bad flag combination) */
#define STB_ERR_CUDA     -2   /* CUDA runtime / no device */
#define STB_ERR_NOMEM    -3   /* device or host allocation failed */
#define STB_ERR_RANGE    -4   /* grid too large for int32 cell index (reference uses int, sidtracemodule.c:32,38) */

/* opaque handle of an acceleration field resident in HBM */
typedef struct stb_field_s *stb_field_t;

/* ---- library ---------------------------------------------------------------- */
int         stb_abi_version(void);
const char *stb_last_error(void);             /* thread-local, never NULL */
int         stb_device_count(int *count);
int         stb_set_device(int device);       /* cudaSetDevice for the calling thread */

/* ---- field: the read-only inputs ax, ay of sidtrace.trace ------------------- */
/* ax, ay: float64, C-contiguous, shape (ny-1, nx-1), row-major with row stride
 * nx-1 (sidtracemodule.c:38; den.py:94-95).  `on_device` != 0 means ax/ay are
 * device pointers (copied device-to-device); otherwise host pointers.  The
 * library keeps its own interleaved copy {ax,ay} per cell (one 16-byte gather
 * per ray-step), so the caller may free ax/ay after the call returns. */
int stb_field_create(const double *ax, const double *ay, int nx, int ny,
                     int on_device, stb_field_t *out);
int stb_field_destroy(stb_field_t f);
int stb_field_shape(stb_field_t f, int *nx, int *ny);
/* read one cell back from HBM: which = 0 -> ax[i, j], 1 -> ay[i, j], indexed
 * with the TRUE row stride nx-1 (the device-side layout pin). */
int stb_field_peek(stb_field_t f, int i, int j, int which, double *out);

/* ---- legacy per-ray entry points -------------------------------------------- */
/* Backs sidtrace.trace (sidtracemodule.c:8-51): x, y, kx, ky are HOST arrays of
 * length L, element 0 = launch state, mutated in place for indices 1..i;
 * *nsteps = i.  One GPU round trip per call: the correctness path, not the fast
 * path. */
int stb_trace(stb_field_t f, double *x, double *y, double *kx, double *ky,
              int64_t L, double st, double dx, double dy, int64_t *nsteps);
/* Backs sidtrace.check (sidtracemodule.c:53-69): *out = a[i*nx + j] on a HOST
 * array (note nx, not nx-1: reproduced as the reference has it). */
int stb_check(int nx, int ny, int i, int j, const double *a, double *out);

/* ---- batched entry point: a whole beam per launch --------------------------- */
/* flags */
#define STB_F_HOST_IO      (1u << 0)  /* x0..ky0, st_ray, amp and all per-ray outputs are HOST pointers
                                         (staged through pinned/HBM workspaces; traj stays a DEVICE pointer) */
#define STB_F_DIV_IEEE     (1u << 1)  /* use the generic IEEE division (div.rn.f64) instead of the
                                         constant-divisor correctly-rounded sequence; both are bit-exact */
#define STB_F_REFILL        (1u << 2)  /* force the persistent grid: warps claim rays from a global queue (a lane
                                         whose ray ended is refilled at once; with STB_F_TRAJ_TILED a whole
                                         32-ray tile at a time) */
#define STB_F_STATIC        (1u << 5)  /* force the static one-thread-per-ray grid.  With neither flag the
                                         library picks: static for tiled trajectories (lock-step rows), lane
                                         refill otherwise -- refill costs <= 15 % on a clean beam but is 3-5x
                                         faster when a few rays run 10-100x longer than the rest (trapped
                                         rays; profiles/r01/kernel_sweep_trapped.txt) */
#define STB_F_NO_HOIST     (1u << 6)  /* keep the 16-byte {ax, ay} cells even when the launch has one global step.
                                         By default such launches read a per-(field, st) table of the four
                                         cell-only products of sidtracemodule.c:41-45 (32 B per cell, built once
                                         and cached on the field; bit-identical results, 14 instead of 20 FP64
                                         instructions per step); tables over 64 MB (STB_HOIST_MAX_MB) are not built */
#define STB_F_YMAJOR       (1u << 7)  /* gather from the y-major copy of the cells (cell (iy, ix) at ix*(ny-1) + iy; built
                                         on first use and cached on the field, like the hoisted tables).  The reference's
                                         a[iy*(nx-1) + ix] (sidtracemodule.c:38) puts the next cell of a ray that climbs in y
                                         one whole row further on -- a new 128-byte line every step; in the y-major copy
                                         it is the neighbouring 16/32 bytes.  Same cells, same bits; set it when the rays
                                         of the launch travel mostly along y (the FT-2 antennas shoot upwards) */
#define STB_F_TRAJ_HOST    (1u << 3)  /* traj is a HOST pointer (copied back after the launch) */
#define STB_F_TRAJ_TILED   (1u << 4)  /* warp-tile schedule: a warp owns 32 consecutive rays until all of them
                                         ended.  traj (if any) is laid out [ray/T][step][ray%T][4] (T =
                                         traj_tile): every warp-step is one coalesced 1 KB row, a CTA writes T/32
                                         adjacent rows (the HBM-friendly layout; traj_ray_stride is ignored).  Per-ray
                                         results leave as coalesced rows too, so with STB_F_HOST_IO and PINNED
                                         host buffers the kernel reads/writes them in place (zero-copy) */

typedef struct stb_batch_s {
    size_t   struct_size;     /* = sizeof(stb_batch_t); ABI evolution guard                     */
    uint32_t flags;
    int64_t  nrays;
    int64_t  L;               /* states per ray incl. the launch state (the reference's MaxI,    */
                              /* sidtracemodule.c:33): at most L-1 steps are taken              */
    /* launch state in GRID units / cm^-1 (ray.py:62-68), one entry per ray */
    const double *x0, *y0, *kx0, *ky0;
    double   st;              /* global step (den.py:107) ...                                    */
    const double *st_ray;     /* ... or a per-ray step (frequency sweeps); NULL -> use `st`      */
    double   dx, dy;          /* n_field.dx[0], n_field.dy[0] (ray.py:72-73)                     */

    /* outputs; any may be NULL */
    double  *traj;            /* [ray][step][4] = (x, y, kx, ky); states 0..i written, rest      */
                              /* untouched (sidtracemodule.c:36-48)                              */
    int64_t  traj_ray_stride; /* in doubles, >= 4*L, multiple of 4 (0 -> 4*L)                    */
    int32_t  traj_tile;       /* STB_F_TRAJ_TILED: rays per tile T = 32 / 64 / 128 / 256 (0 -> 256):    */
                              /* layout [ray/T][step][ray%T][4], ceil(nrays/T)*L*T*4 doubles; T = 256 = */
                              /* one 256-thread CTA = 8 KB contiguous rows (0.98 of the HBM copy peak)  */
    int32_t *nsteps;          /* [ray] i = steps taken = index of last written state            */
    double  *end_state;       /* [4][nrays] state at index i                                     */

    /* fused per-ray reductions (ray.py:113-126 and SURVEY 8c extensions); may be NULL */
    int32_t *turn_index;      /* [ray] first index j in [1,i] where ky changes sign, else -1     */
    double  *turn_state;      /* [4][nrays] (x, y, kx, ky) of state turn_index (grid units), NaN  */
                              /* when the ray has no turning point                               */
    double  *lax_end;         /* [ray] lax[i-1]: path-length integral over the kept states       */
    double  *ph_end;          /* [ray] ph[i-1]:  eikonal phase over the kept states              */

    /* deposition-profile extension (no reference counterpart; oracle/sidtrace_oracle.c) */
    double  *hist;            /* [nbins] accumulated (not cleared); DEVICE pointer unless HOST_IO */
    int32_t  nbins;
    const double *amp;        /* [ray] launch amplitude (ray.py:56); NULL -> 1                   */
    double   xorg, yorg;      /* physical coordinates of grid node (0,0)                         */
    double   r_lcs, rho_max;

    /* receiver-line crossing, fused: what the reference's Ray.addpoint(val, axis)
     * (ray.py:131-224) reports for the FIRST crossing of the line y = cross_val (axis 1) or
     * x = cross_val (axis 2), in PHYSICAL cm -- the inputs of Beam.field_r (beam.py:359-369).
     * cross_axis = 0 disables it.  Tolerance-level parity (1e-9), see oracle/crossing.py. */
    int32_t  cross_axis;
    double   cross_val;
    const double *x0_phys, *y0_phys;   /* [ray] physical launch point (ray.py:105-106); required */
    int32_t *cross_count;     /* [ray] crossing segments found (0: the reference returns nan)      */
    int32_t *cross_index;     /* [ray] index addpoint returns for the first crossing, -1 if none   */
    int32_t *cross_flags;     /* [ray] bit 0: exact hit of an existing point; bit 1: extrapolated  */
                              /*       past the last kept point (ray.py:147-148)                   */
    double  *cross_state;     /* [6][nrays] x, y (cm), kx, ky, lax, ph at that point; NaN if none  */

    /* equal-phase lookup (needs cross_axis != 0): the neighbour-ray step of Beam.field_r
     * (beam.py:404-425).  Each ray -- as it stands after addpoint inserted its crossing point and
     * its phase was recomputed (ray.py:205-216, 122-126) -- is sampled at two target phases with
     * the linear interpolation / end-segment extrapolation of scipy's interp1d.  The rays are
     * traced again; no stored trajectory is needed. */
    const double *lookup_ph;  /* [2][nrays] target phases A, B per ray; NULL disables             */
    double  *lookup_xy;       /* [4][nrays] xA, yA, xB, yB (cm); NaN if the ray has < 2 points     */

    /* packed tiled trajectories (with STB_F_TRAJ_TILED; NULL -> the [ray/T][step][ray%T][4] layout with
     * L rows per tile): the plan stb_traj_plan made from the step counts of a first pass.  Slot s of
     * the plan holds ray traj_perm[s] (-1: empty); tile t = slots [t*T, (t+1)*T) owns rows
     * traj_tile_off[t] .. traj_tile_off[t+1] of T x 4 doubles; state j of the ray in slot s is at
     * traj[((traj_tile_off[s/T] + j)*T + s%T)*4].  No ray steps past the rows of its tile.
     * traj_perm WITHOUT traj_tile_off (any flags, device I/O): the ORDER the rays are traced in -- slot s is ray
     * traj_perm[s] (-1: empty), traj_nslots >= nrays slots.  For launch sets that arrive in random order (sweeps):
     * sorted by launch cell the lanes of a warp gather from the same few cells.  Per-ray outputs and ray-major
     * trajectories are written at the ray's own index; tiled trajectories are laid out by SLOT
     * (traj_nslots = ceil(nrays/T)*T). */
    const int32_t *traj_perm;      /* [traj_nslots]                                              */
    const int64_t *traj_tile_off;  /* [traj_nslots/T + 1]                                        */
    int64_t  traj_nslots;          /* ceil(nrays/T)*T                                            */
} stb_batch_t;

/* Traces b->nrays rays on `stream` (a cudaStream_t passed as void*; NULL = the
 * legacy default stream).  With device pointers the call is asynchronous;
 * with STB_F_HOST_IO / STB_F_TRAJ_HOST it returns after the results are in
 * host memory. */
int stb_trace_batch(stb_field_t f, const stb_batch_t *b, void *stream);

/* ---- launch states on the device (ray.py:34-68 for a whole beam) ------------------------------- */
/* From physical start points and directions to what stb_trace_batch takes: direction normalised to
 * w/c and scaled by the local refractive index N = sqrt(1 - n/n_c) (n from the bicubic B-spline
 * (tx, ty, coef) SciPy fitted to the density grid, evaluated like FITPACK fpbisp), start points
 * mapped to grid units with the piecewise-linear x2grid / y2grid of den.py:69-70 (bit-identical to
 * scipy's interp1d).  All pointers are DEVICE pointers; asynchronous on `stream`. */
typedef struct stb_launch_s {
    size_t   struct_size;
    int64_t  n;
    const double *x0_phys, *y0_phys;   /* [n] start points, cm                                    */
    const double *kx_dir, *ky_dir;     /* [n] direction (any non-zero length)                     */
    const double *x_nodes, *y_nodes;   /* [nx], [ny] grid node coordinates (Nfield.x, Nfield.y)   */
    int32_t  nx, ny;
    const double *tx, *ty, *coef;      /* knots [ntx], [nty] and coefficients [(ntx-4)*(nty-4)]   */
    int32_t  ntx, nty;
    double   w, c, n_crit;             /* angular frequency, speed of light (cm/s), critical density */
    double  *x0, *y0, *kx0, *ky0;      /* [n] outputs: grid units / cm^-1                         */
} stb_launch_t;
int stb_launch_states(const stb_launch_t *a, void *stream);

/* Host helper (no device work): the "pages" of Beam.field_r (beam.py:374-398) -- the rays, in launch
 * order with receiver coordinate xout[n], split into runs of increasing xout; pages of fewer than two
 * rays are dropped.  members needs room for 2*n entries, offsets for n+2; page k is
 * members[offsets[k] .. offsets[k+1]). */
int stb_field_pages(const double *xout, int64_t n, int64_t *offsets, int64_t *members, int64_t *npages);

/* ---- the O(rays) tail of Beam.field_r / field_result on the device --------------------- */
/* The pages again, as per-ray links and in O(n log P) (host, no device work): pred[j] / succ[j] = the ray
 * before / after ray j in its page (-1 at the page's ends), page_size[j] = rays in ray j's page (pages of
 * one ray are dropped by the reference, beam.py:392-398), *npages = pages of two or more rays.  The walk of
 * beam.py:374-398 is first-fit -- a ray joins the first page whose last member it lies beyond -- and the
 * pages' last members are non-increasing in the page index, so each ray is placed by bisection. */
int stb_field_pages_links(const double *xout, int64_t n, int32_t *pred, int32_t *succ, int32_t *page_size, int64_t *npages);

/* Ray-tube amplitude and complex field on the receiver line (beam.py:404-427) for the n rays that cross it
 * exactly once, in launch order: amp1 = sqrt((d0l + d0r)/(d1l + d1r)) with d0 the distance to the page
 * neighbours at launch and d1 at equal phase (lookup_xy = the lookup pass' output for targets A = phase of
 * succ, B = phase of pred), NaN -> amp0; field = amp0 * amp1 * exp(-i ph), 0 for rays of dropped pages.
 * DEVICE pointers; field is interleaved complex [n][2]. */
typedef struct stb_tube_s {
    size_t   struct_size;
    int64_t  n;
    const int32_t *pred, *succ, *page_size;
    const double *x0, *y0;          /* launch points, cm                                   */
    const double *xc, *yc, *ph;     /* crossing point and eikonal phase (cross_state 0, 1, 5) */
    const double *amp0;
    const double *lookup_xy;        /* [4][n]                                              */
    double  *amp1;                  /* [n] out                                             */
    double  *field;                 /* [n][2] out                                          */
} stb_tube_t;
int stb_tube_field(const stb_tube_t *a, void *stream);
/* beam.py:436-440 on the arrays sorted by receiver coordinate: rays whose field is 0 take the mean of their
 * two neighbours' amplitude (np.roll: cyclic), NaN -> amp0, and field = amp0 * amp1 * exp(-i ph).
 * amp1_in != amp1_out (the neighbours' OLD values are read). */
int stb_hole_fill(int64_t n, const double *amp0, const double *ph, const double *amp1_in, double *amp1_out, double *field, void *stream);

/* field_result, first half (beam.py:448-461): field = field_tx x the analytic pattern of the RX beam at
 * (x, y_rx) (beam.py:103-131, 176-206, n = 1); sums[0..4] = sum|f|, sum|f| x, Re / Im of sum f dist
 * (dist = centred spacing of x, beam.py:455-457), max|f|.  sums has room for 8 doubles. */
typedef struct stb_overlap_s {
    size_t   struct_size;
    int64_t  n;
    const double *x;                /* [n] receiver coordinate, sorted                     */
    const double *field_tx;         /* [n][2]                                              */
    double   y_rx;
    double   bx0, by0, bkx, bky, focus, waist, wave_len, rayleigh, gouy;   /* RX beam      */
    double  *field;                 /* [n][2] out                                          */
    double  *absf;                  /* [n] out                                             */
    double  *sums;                  /* [8] out (zeroed by the call)                        */
} stb_overlap_t;
int stb_overlap(const stb_overlap_t *a, void *stream);
/* field_result, second half (beam.py:463-467): what one Levenberg-Marquardt iteration of the fit
 * y ~ amp exp(-(x - xc)^2 / (2 sigma^2)) needs at the current parameters: out[0..5] = J^T J (aa ab ac bb bc cc),
 * out[6..8] = J^T r, out[9] = r^T r (r = model - y).  fd_step (HOST, 3 doubles) != NULL: the Jacobian columns
 * are MINPACK's forward differences with these steps (what scipy.optimize.curve_fit uses without jac=);
 * NULL: analytic columns.  out (DEVICE, 10 doubles) is zeroed by the call. */
int stb_gauss_fit_sums(const double *x, const double *y, int64_t n, double amp, double xc, double sigma, const double *fd_step,
                       double *out, void *stream);

/* ---- beam launch conditions on the device (Beam.rays, beam.py:293-331) -------------------------- */
/* n start points equidistant in arc length along a wavefront given by nw nodes (arc length lax[],
 * beam-frame x wf_x[]) and the cubic B-spline (t, c) of its y(x); unit normals from the reference's
 * central differences on the lcm(n, nfront)-point fine grid; Gaussian amplitudes.  Outputs are lab-
 * frame start points, unit directions and amplitudes -- the inputs of stb_launch_states.
 * All pointers are DEVICE pointers; asynchronous on `stream`.  Tolerance-level (1e-12) agreement with
 * the host restatement (SciPy evaluates the same spline). */
typedef struct stb_beam_s {
    size_t   struct_size;
    int64_t  n;                 /* rays (rnum)                                              */
    int32_t  nfront;            /* Beam._nfront (1000): the fine grid has lcm(n, nfront) points */
    const double *lax, *wf_x;   /* [nw]                                                      */
    int32_t  nw;
    const double *t, *c;        /* knots [nt] and coefficients [nt-4] of the cubic y(x)      */
    int32_t  nt;
    double   x0, y0, kx0, ky0, focus;   /* antenna position, unit direction, focus offset    */
    double   waist, rayleigh;           /* w0 and the Rayleigh length at the edge index      */
    double  *x, *y, *kx, *ky, *amp;     /* [n] outputs                                       */
} stb_beam_t;
int stb_beam_launch(const stb_beam_t *a, void *stream);

/* ---- exactly-sized trajectory arenas -------------------------------------------------------------- */
/* The reference gives every ray MaxStep-long arrays and keeps [:i] (ray.py:19-22, 99-102).  Here a first
 * pass without trajectory output (step counts only) is followed by this plan: rays sorted by step count,
 * longest first, ties in launch order (stable radix sort); T = `tile` consecutive ones share a tile; a tile
 * owns (step count of its longest ray + 1) rows.  Outputs (DEVICE): perm[ceil(nrays/T)*T] slot -> ray
 * (-1 = empty), slot_of[nrays] ray -> slot, tile_off[ceil(nrays/T)+1] first row of every tile; *total_rows
 * (HOST) = rows of the whole arena, i.e. it needs total_rows*T*32 bytes.  L > 0 tells the sort how many
 * key bits matter (nsteps < L).  Synchronises `stream` before returning. */
int stb_traj_plan(const int32_t *nsteps, int64_t nrays, int32_t tile, int64_t L,
                  int32_t *perm, int32_t *slot_of, int64_t *tile_off, int64_t *total_rows, void *stream);

/* Trajectory export: packs the KEPT states 0..nsteps[r]-1 of every ray (what the reference's Ray keeps,
 * ray.py:99-102) from a trajectory arena -- ray-major (tile = 0, ray_stride doubles per ray, 0 -> 4*L) or
 * tiled (tile = 32 / 64 / 128) -- into out[offsets[r] .. offsets[r] + nsteps[r]][4].  All pointers are
 * DEVICE pointers; offsets is an exclusive prefix sum of nsteps (int64).  Asynchronous on `stream`. */
int stb_pack_traj(const double *traj, int tile, int64_t ray_stride, int64_t L, int64_t nrays,
                  const int32_t *nsteps, const int64_t *offsets, double *out, void *stream);
/* The same from a packed tiled arena (stb_traj_plan): slot_of, tile_off as the plan wrote them. */
int stb_pack_traj_plan(const double *traj, int tile, const int32_t *slot_of, const int64_t *tile_off, int64_t nrays,
                       const int32_t *nsteps, const int64_t *offsets, double *out, void *stream);

/* number of kernel launches issued by this library so far (all threads) */
int64_t stb_launch_count(void);

/* ---- measurement helpers ----------------------------------------------------- */
/* Dependent-free DFMA throughput of the current device in TFLOP/s (2 flops per
 * DFMA) -- the FP64 roofline denominator that MEASURED_PEAKS.json lacks. */
int stb_measure_fp64_peak(double *tflops);
/* Plain streaming-store bandwidth in GB/s over `bytes` of HBM. */
int stb_measure_store_bw(size_t bytes, double *gbs);
/* The store pattern of the static tiled trajectory kernel without its arithmetic: `ntiles` CTAs of `tile`
 * threads, each thread writes `steps` 32-byte states at stride tile*32 B into its tile of `rows_per_tile`
 * rows, with `chain` dependent DFMAs between two stores.  GB/s of the best of 4 timed launches. */
int stb_measure_store_pattern(int64_t ntiles, int32_t tile, int64_t rows_per_tile, int32_t steps, int32_t chain, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* SIDTRACE_B200_H */
