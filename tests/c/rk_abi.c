/* A plain-C caller of the adaptive-RK extension's C ABI (include/sidtrace_b200_rk.h) with host
 * pointers (STB_RK_F_HOST_IO): closed-form checks only -- the reference has no counterpart of this
 * mode.  Vacuum: straight line, optical path = tau.  Linear ramp X = a + b x, O mode:
 * x(T) = x0 + px T - b T^2/4, px(T) = px - b T/2.  Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/sidtrace_b200_rk.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != STB_OK) { \
    fprintf(stderr, "%s -> %d: %s\n", #call, rc_, stb_last_error()); return 2; } } while (0)

enum { N = 101, NR = 5 };

static int run(stb_medium_t m, const double *x0, const double *y0, const double *px, const double *py, double T,
               double *end, double *phi, int32_t *status, int32_t *nacc)
{
    stb_rk_t k;
    memset(&k, 0, sizeof k);
    k.struct_size = sizeof k;
    k.flags = STB_RK_F_HOST_IO;
    k.nrays = NR;
    k.x0 = x0; k.y0 = y0; k.px0 = px; k.py0 = py;
    k.rtol = 1e-10; k.atol_r = 1e-10; k.atol_p = 1e-10;
    k.h0 = 0.1; k.hmin = 1e-12; k.hmax = 0.5; k.tau_max = T; k.max_steps = 100000;
    k.end_state = end; k.phi_end = phi; k.status = status; k.nacc = nacc;
    return stb_trace_rk(m, &k, NULL);
}

int main(void)
{
    static double X[N * N];
    const double a = 0.4, b = 0.04, org = -10.0, d = 0.2;
    stb_medium_desc_t desc;
    memset(&desc, 0, sizeof desc);
    desc.struct_size = sizeof desc;
    desc.nx = N; desc.ny = N; desc.xorg = org; desc.yorg = org; desc.dx = d; desc.dy = d;
    desc.X = X; desc.mode = STB_RK_MODE_O; desc.interp = STB_RK_INTERP_BSPLINE;
    int bad = 0;
    double x0[NR], y0[NR], px[NR], py[NR], end[4 * NR], phi[NR];
    int32_t status[NR], nacc[NR];

    /* vacuum */
    stb_medium_t vac;
    CHECK(stb_medium_create(&desc, &vac));
    for (int r = 0; r < NR; ++r) { x0[r] = -9.0; y0[r] = -9.0 + r; px[r] = cos(0.1 * r); py[r] = sin(0.1 * r); }
    CHECK(run(vac, x0, y0, px, py, 11.0, end, phi, status, nacc));
    for (int r = 0; r < NR; ++r) {
        bad |= status[r] != STB_RK_TAU_MAX;
        bad |= fabs(end[r] - (x0[r] + 11.0 * px[r])) > 1e-13 || fabs(end[NR + r] - (y0[r] + 11.0 * py[r])) > 1e-13;
        bad |= fabs(phi[r] - 11.0) > 1e-13;
    }
    printf("vacuum: end.x=%.15g phi=%.15g steps=%d %s\n", end[0], phi[0], nacc[0], bad ? "MISMATCH" : "ok");
    CHECK(stb_medium_destroy(vac));

    /* linear ramp */
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) X[j * N + i] = a + b * (org + d * i);
    stb_medium_t ramp;
    CHECK(stb_medium_create(&desc, &ramp));
    const double T = 6.0, N0 = sqrt(1.0 - (a + b * -9.5));
    for (int r = 0; r < NR; ++r) { x0[r] = -9.5; y0[r] = -2.0 + r; px[r] = N0 * cos(0.2 * r - 0.4); py[r] = N0 * sin(0.2 * r - 0.4); }
    CHECK(run(ramp, x0, y0, px, py, T, end, phi, status, nacc));
    for (int r = 0; r < NR; ++r) {
        bad |= status[r] != STB_RK_TAU_MAX;
        bad |= fabs(end[r] - (x0[r] + px[r] * T - 0.25 * b * T * T)) > 5e-12;
        bad |= fabs(end[NR + r] - (y0[r] + py[r] * T)) > 5e-12;
        bad |= fabs(end[2 * NR + r] - (px[r] - 0.5 * b * T)) > 5e-12 || fabs(end[3 * NR + r] - py[r]) > 5e-12;
    }
    printf("ramp: end.x=%.15g (closed form %.15g) %s\n", end[0], x0[0] + px[0] * T - 0.25 * b * T * T, bad ? "MISMATCH" : "ok");

    /* the spline table the kernel reads interpolates the nodes: (c[i-1] + 4 c[i] + c[i+1]) / 6 = X[i] */
    static double coef[(N + 2) * (N + 2)];
    CHECK(stb_medium_table(ramp, 0, coef, (size_t)(N + 2) * (N + 2)));
    const double *c = coef + 51 * (N + 2) + 51;   /* coefficient (50, 50) */
    const double row_m = (c[-(N + 2) - 1] + 4 * c[-(N + 2)] + c[-(N + 2) + 1]) / 6, row_0 = (c[-1] + 4 * c[0] + c[1]) / 6,
                 row_p = (c[(N + 2) - 1] + 4 * c[N + 2] + c[(N + 2) + 1]) / 6;
    bad |= fabs((row_m + 4 * row_0 + row_p) / 6 - X[50 * N + 50]) > 1e-15;

    /* error convention */
    int rc = stb_medium_table(ramp, 1, coef, 1);
    bad |= rc != STB_ERR_ARG || strlen(stb_last_error()) == 0;
    desc.interp = 7;
    stb_medium_t none = NULL;
    rc = stb_medium_create(&desc, &none);
    bad |= rc != STB_ERR_ARG || none != NULL;
    CHECK(stb_medium_destroy(ramp));
    printf("%s\n", bad ? "MISMATCH" : "all ok");
    return bad ? 1 : 0;
}
