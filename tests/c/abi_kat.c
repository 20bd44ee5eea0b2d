/* A plain-C caller of the C ABI (include/sidtrace_b200.h): the known-answer vector G1 of
 * SURVEY.md 4.1 through stb_trace (legacy per-ray entry) and stb_trace_batch (device buffers are
 * not needed: STB_F_HOST_IO stages host pointers).  No Python, no torch: this is what a non-Python
 * host binds.  Exit code 0 = bit-identical to the reference's published values. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/sidtrace_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != STB_OK) { \
    fprintf(stderr, "%s -> %d: %s\n", #call, rc_, stb_last_error()); return 2; } } while (0)

int main(void)
{
    if (stb_abi_version() != STB_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }
    /* G1: nx=11, ny=9, ax=0.5, ay=-0.25 on the (8, 10) cell grid, st=0.5, dx=1, dy=2 */
    enum { NX = 11, NY = 9, L = 4096 };
    static double ax[(NY - 1) * (NX - 1)], ay[(NY - 1) * (NX - 1)];
    for (int i = 0; i < (NY - 1) * (NX - 1); ++i) { ax[i] = 0.5; ay[i] = -0.25; }
    stb_field_t f;
    CHECK(stb_field_create(ax, ay, NX, NY, 0, &f));
    double *x = calloc(L, 8), *y = calloc(L, 8), *kx = calloc(L, 8), *ky = calloc(L, 8);
    x[0] = 1.2; y[0] = 3.7; kx[0] = 0.3; ky[0] = 0.9;
    int64_t n = -1;
    CHECK(stb_trace(f, x, y, kx, ky, L, 0.5, 1.0, 2.0, &n));
    int bad = 0;
    bad |= n != 11;
    bad |= x[11] != 0x1.4d33333333334p+3;      /* 10.412500000000001 */
    bad |= y[11] != 0x1.1233333333330p+2;      /*  4.284374999999997 */
    bad |= fabs(kx[11] - 3.05) > 1e-15 || fabs(ky[11] + 0.475) > 1e-15;
    bad |= x[12] != 0.0;                       /* untouched beyond state n */
    printf("stb_trace: n=%lld x[11]=%a y[11]=%a %s\n", (long long)n, x[11], y[11], bad ? "MISMATCH" : "ok");

    /* the same ray three times through the batched entry with host pointers */
    double x0[3] = {1.2, 1.2, -0.5}, y0[3] = {3.7, 3.7, 3.7}, kx0[3] = {0.3, 0.3, 0.3}, ky0[3] = {0.9, 0.9, 0.9};
    int32_t ns[3] = {-1, -1, -1};
    double end[12];
    stb_batch_t b;
    memset(&b, 0, sizeof b);
    b.struct_size = sizeof b;
    b.flags = STB_F_HOST_IO;
    b.nrays = 3; b.L = L;
    b.x0 = x0; b.y0 = y0; b.kx0 = kx0; b.ky0 = ky0;
    b.st = 0.5; b.dx = 1.0; b.dy = 2.0;
    b.nsteps = ns; b.end_state = end;
    CHECK(stb_trace_batch(f, &b, NULL));
    bad |= ns[0] != 11 || ns[1] != 11 || ns[2] != 0;          /* G6: a start outside the grid returns 0 */
    bad |= end[0] != x[11] || end[3 + 0] != y[11] || end[1] != x[11];
    printf("stb_trace_batch: nsteps=%d,%d,%d end.x=%a %s\n", ns[0], ns[1], ns[2], end[0], bad ? "MISMATCH" : "ok");

    double v = 0;
    CHECK(stb_check(NX - 1, NY - 1, 2, 3, ax, &v));
    bad |= v != 0.5;
    CHECK(stb_field_peek(f, 7, 9, 1, &v));
    bad |= v != -0.25;
    /* error convention: status + message, nothing thrown */
    b.L = 0;
    int rc = stb_trace_batch(f, &b, NULL);
    bad |= rc != STB_ERR_ARG || strlen(stb_last_error()) == 0;
    CHECK(stb_field_destroy(f));
    free(x); free(y); free(kx); free(ky);
    return bad ? 1 : 0;
}
