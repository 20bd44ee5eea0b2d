/* sidtrace_b200_rk.h -- C ABI of the beyond-reference physics mode.
 *
 * EXTENSION, NOT PARITY.  The reference's only integrator is the fixed-step loop of
 * sidtracemodule.c:36-48 on nearest-cell accelerations (den.py:84-107); that path is
 * include/sidtrace_b200.h.  This header adds what BASELINE.json's north_star describes beyond it:
 * an adaptive embedded Runge-Kutta (Dormand-Prince 5(4)) integration of the Hamiltonian ray
 * equations
 *
 *     dr/dtau = p,    dp/dtau = (1/2) grad N^2(r),    p = k/k0,    H = |p|^2 - N^2(r)
 *
 * through an INTERPOLATED medium (cubic-spline, Catmull-Rom or bilinear n/n_c table -- the table the
 * reference differences in den.py:65-66 and splines for its launch index, den.py:71 -- with the cold-plasma O- or X-mode dispersion and its
 * analytic derivatives fused into the evaluation; the X mode's field is the analytic toroidal B ~ 1/R with the
 * rmain/btor of equilibrium.py:289, or |B| of a tabulated psi(R, Z) + F(psi) in the format of equilibrium.py:11-49),
 * optical-depth integration of an absorption table, and the absorbed
 * power deposited into flux-coordinate bins.  There is no reference output to compare with: the
 * checker is oracle/rk_oracle.c (same operation order, bit-identical) pinned by analytic
 * solutions and an independent SciPy integration (tests/test_rk_*.py).
 *
 * Same conventions as sidtrace_b200.h: 0 = success, negative = error, stb_last_error().
 * Array pointers are DEVICE pointers unless STB_RK_F_HOST_IO is set.
 */
#ifndef SIDTRACE_B200_RK_H
#define SIDTRACE_B200_RK_H

#include "sidtrace_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stb_medium_s *stb_medium_t;

/* interpolation of the node tables */
#define STB_RK_INTERP_BILINEAR    0   /* C0: force jumps at cell edges                                  */
#define STB_RK_INTERP_CATMULL_ROM 1   /* C1 bicubic patches straight from the nodes (local, no prefilter) */
#define STB_RK_INTERP_BSPLINE     2   /* C2 natural cubic spline: the library prefilters the tables on  */
                                      /* the device into (ny+2) x (nx+2) B-spline coefficients.  The    */
                                      /* smooth choice: the step controller sees no kinks               */

#define STB_RK_MODE_O 0   /* N^2 = 1 - X                                   */
#define STB_RK_MODE_X 1   /* N^2 = 1 - X(1-X)/(1-X-Y^2), Y = Y0 R0/(R0+x)  */
#define STB_RK_MODE_X_TABLE 2   /* the same dispersion with Y^2 = (omega_ce/omega)^2 interpolated from a node table on
                                   the medium's grid (like X): |B| of a TABULATED equilibrium -- given directly (Y2) or
                                   built by the library on the device from psi(R, Z) and F(psi) (equilibrium) */

/* per-ray end status */
#define STB_RK_LEFT_DOMAIN    0   /* accepted state outside the table (the reference's loop exit, sidtracemodule.c:36) */
#define STB_RK_TAU_MAX        1
#define STB_RK_MAX_STEPS      2
#define STB_RK_STEP_UNDERFLOW 3   /* step shrank below hmin: resonance (N^2 -> inf) or NaN input */
#define STB_RK_START_OUTSIDE  4
#define STB_RK_STOP_LINE      5   /* landed on the stop line (receiver line, cf. Ray.addpoint, ray.py:131-224) */

/* A tabulated axisymmetric equilibrium in the geqdsk-like form of the reference's Equilibrium class
 * (equilibrium.py:11-49: psi on an nr x nz rectangle [rleft, rleft+rdim] x [zmid-zdim/2, zmid+zdim/2], R-major
 * "so (x,y) axis order is in geqdsk format", interpolated BILINEARLY -- RectBivariateSpline(kx=1, ky=1),
 * equilibrium.py:46-49 -- and fpol = R B_phi on a uniform psi grid from simagx to sibdry).  SI units (m, Wb/rad, T m).
 * From it, at every node (x, y) of the medium's grid (R = (r0_cm + x)/100, Z = (y - z0_cm)/100):
 *     B_R = -(1/R) dpsi/dZ,  B_Z = (1/R) dpsi/dR,  B_phi = F(psi)/R  (F held at fpol[nf-1] outside the last surface)
 *     (psi itself bilinear; grad psi = central differences at the table's nodes, interpolated bilinearly, so that the
 *      poloidal field is continuous -- the gradient of the bilinear patch jumps at every cell edge)
 *     Y^2 = (y_per_tesla)^2 (B_R^2 + B_Z^2 + B_phi^2),   y_per_tesla = e/(m_e omega)
 * No reference code evaluates B from this table (equilibrium.py is never called by the tracer); this is the
 * north_star's "psi(R, Z), B components" for the extension's X mode. */
typedef struct stb_equilibrium_s {
    size_t struct_size;
    int32_t nr, nz;
    double rleft, rdim, zmid, zdim;   /* m */
    const double *psi;                /* [nr][nz], HOST */
    const double *fpol;               /* [nf], HOST */
    int32_t nf;
    double simagx, sibdry;
    double r0_cm, z0_cm;              /* where the medium's (x, y) = (0, 0) sits: R = r0_cm, Z = -z0_cm ... see above */
    double y_per_tesla;
} stb_equilibrium_t;

typedef struct stb_medium_desc_s {
    size_t struct_size;
    int32_t nx, ny;            /* nodes; the table is [ny][nx] row-major like den.py's n            */
    double xorg, yorg;         /* physical position (cm) of node (0, 0)                             */
    double dx, dy;             /* node spacing (cm)                                                 */
    const double *X;           /* n/n_c at the reference frequency (den.py:335-348 gives n_c)       */
    const double *alpha;       /* power absorption coefficient per unit tau (1/cm), or NULL         */
    int32_t on_device;         /* 0: X/alpha are host pointers (copied), 1: device pointers (copied)*/
    int32_t mode;              /* STB_RK_MODE_*                                                     */
    int32_t interp;            /* STB_RK_INTERP_*                                                   */
    double Y0, R0;             /* X mode: omega_ce/omega on the axis x = 0, major radius (cm)       */
    const double *Y2;          /* STB_RK_MODE_X_TABLE: (omega_ce/omega)^2 on the nodes, [ny][nx], where X lives;   */
                               /* or NULL and ...                                                                  */
    const stb_equilibrium_t *equilibrium;   /* ... the tabulated equilibrium the library builds that table from     */
} stb_medium_desc_t;

int stb_medium_create(const stb_medium_desc_t *desc, stb_medium_t *out);
int stb_medium_destroy(stb_medium_t m);
/* Copies the table the kernel reads back to the host: which = 0 -> X, 1 -> alpha, 2 -> Y^2; [ny][nx] doubles, or
 * the [ny+2][nx+2] spline coefficients for STB_RK_INTERP_BSPLINE (count must match). */
int stb_medium_table(stb_medium_t m, int which, double *host_out, size_t count);

#define STB_RK_F_LANE_REFILL (1u << 1) /* same as refill_min = 1 */
#define STB_RK_F_HOST_IO (1u << 0)   /* every array pointer of stb_rk_t is a HOST pointer: the call stages them
                                        through HBM and returns when the results are back (synchronous) */

typedef struct stb_rk_s {
    size_t   struct_size;
    uint32_t flags;            /* STB_RK_F_*                                                        */
    int64_t  nrays;
    const double *x0, *y0;     /* launch position, cm                                               */
    const double *px0, *py0;   /* launch refractive-index vector k/k0 (|p| = N on the ray H = 0)    */
    const double *xscale;      /* per-ray multiplier of X (density / frequency sweeps) or NULL      */
    const double *yscale;      /* per-ray multiplier of Y or NULL                                   */
    const double *amp;         /* launch amplitude or NULL (-> 1); power = amp^2                    */
    double rtol, atol_r, atol_p;   /* error control: |e_i| <= atol + rtol*|y_i| (max norm)          */
    double h0, hmin, hmax;     /* first, smallest and largest step in tau (cm)                      */
    double tau_max;            /* stop at this tau (HUGE_VAL: never)                                */
    int32_t max_steps;         /* accepted-step cap                                                 */
    int32_t refill_min;        /* schedule: a warp claims new rays (consecutive indices, one atomic) once this many of its */
                               /* lanes are idle.  32 = "warp tiles": the warp owns 32 consecutive rays at a time --       */
                               /* neighbouring rays of a beam share cells, so its stencil loads coalesce (fastest on a     */
                               /* clean, ordered beam: 3.7 ms per 10^6 rays); 1 = lane refill (4.2 ms; best for random ray  */
                               /* order or when a few rays take 10-100x the steps of the rest); 0 -> 24, which costs 4 %   */
                               /* on the clean beam and never leaves more than 23 lanes waiting for a straggler            */
    int32_t stop_axis;         /* 0: none; 1: end the ray ON the line y = stop_val; 2: on x = stop_val. The step  */
                               /* that crosses the line is repeated with shorter lengths (regula falsi on the   */
                               /* step length, counted in nrej) until |coordinate - stop_val| <= stop_tol         */
    double   stop_val, stop_tol;

    double  *end_state;        /* [4][nrays] (x, y, px, py)                                         */
    double  *tau_end;          /* [ray] integration variable at the end (path length in vacuum)     */
    double  *phi_end;          /* [ray] optical path  int |p|^2 dtau  (phase = k0 * phi)            */
    double  *sig_end;          /* [ray] optical depth int alpha dtau                                */
    double  *n2_min;           /* [ray] smallest N^2 over the accepted states (depth of penetration)*/
    double  *ham_end;          /* [ray] H = |p|^2 - N^2 at the end state (conservation diagnostic)  */
    int32_t *nacc, *nrej;      /* [ray] accepted / rejected steps                                   */
    int32_t *status;           /* [ray] STB_RK_*                                                    */
    int32_t *turn_index;       /* [ray] first accepted step at which p_y changed sign (turning point, the  */
                               /* definition of stb_batch_t.turn_index), or -1                             */
    double  *turn_state;       /* [4][nrays] (x, y, px, py) there; NaN when the ray has no turning point    */
    double  *traj;             /* [ray][L][4] accepted states, state 0 = launch; states >= L dropped*/
    double  *traj_aux;         /* [ray][L][2] (tau, sigma) of the same states                       */
    int64_t  L;

    double  *hist;             /* [nbins] absorbed power per rho bin, accumulated (needs alpha)     */
    int32_t  nbins;
    double   xaxis, yaxis;     /* magnetic axis (cm)                                                */
    double   r_lcs, rho_max;   /* rho = |r - axis| / r_lcs, bins cover [0, rho_max)                 */
} stb_rk_t;

/* Traces all rays of `rk` through `m` on `stream` (a cudaStream_t, may be NULL).  Asynchronous unless
 * STB_RK_F_HOST_IO is set. */
int stb_trace_rk(stb_medium_t m, const stb_rk_t *rk, void *stream);

#ifdef __cplusplus
}
#endif
#endif
