/* A plain-C, two-process caller of the multi-GPU entry points (include/sidtrace_b200_comm.h): no Python,
 * no torch.  Rank 0 draws the NCCL unique id and hands it to rank 1 through a pipe; each rank, on its own
 * GPU, traces its block-interleaved shard of ONE beam (stb_trace_batch, device buffers, deposition histogram),
 * the histogram is all-reduced, the per-ray step counts and end states are gathered on rank 0 and put
 * back in global ray order (stb_comm_gather + stb_unshuffle_blocks).  Rank 0 then traces the whole beam
 * alone and compares: step counts and end states bit for bit, histogram to 1e-12 (summation order).
 *
 *   gcc -O2 comm_2rank.c -I<repo>/include -I/usr/local/cuda/include -L<lib dir> -lsidtrace_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -o comm_2rank        (run under gpurun --gpus 2)
 * Exit code 0 = identical; 77 = fewer than two GPUs (skipped). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cuda_runtime_api.h>
#include "../../include/sidtrace_b200_comm.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != STB_OK) { \
    fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc_, stb_last_error()); return 2; } } while (0)
#define CUCHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "rank %d: %s: %s\n", rank, #call, cudaGetErrorString(e_)); return 2; } } while (0)

enum { NX = 65, NY = 49, N = 5000, B = 256, WORLD = 2, L = 400, NBINS = 16 };

static double frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (*s >> 8) / 16777216.0; }

static int trace(int rank, stb_field_t f, const double *h_launch /* [4][n] */, int64_t n, int32_t *d_ns, double *d_end,
                 double *d_hist)
{
    double *d_in = NULL;
    CUCHECK(cudaMalloc((void **)&d_in, (size_t)(4 * n + 1) * sizeof(double)));
    CUCHECK(cudaMemcpy(d_in, h_launch, (size_t)4 * n * sizeof(double), cudaMemcpyHostToDevice));
    stb_batch_t b;
    memset(&b, 0, sizeof b);
    b.struct_size = sizeof b;
    b.flags = STB_F_STATIC;
    b.nrays = n; b.L = L;
    b.x0 = d_in; b.y0 = d_in + n; b.kx0 = d_in + 2 * n; b.ky0 = d_in + 3 * n;
    b.st = 0.3; b.dx = 1.0; b.dy = 0.8;
    b.nsteps = d_ns; b.end_state = d_end;
    b.hist = d_hist; b.nbins = NBINS; b.xorg = -32.0; b.yorg = -20.0; b.r_lcs = 25.0; b.rho_max = 1.5;
    CHECK(stb_trace_batch(f, &b, NULL));
    CUCHECK(cudaDeviceSynchronize());
    cudaFree(d_in);
    return 0;
}

static int run(int rank, int fd_in, int fd_out)
{
    int ndev = 0;
    CHECK(stb_device_count(&ndev));
    if (ndev < WORLD) { if (rank == 0) printf("comm_2rank: %d GPU(s) visible, need %d -- skipped\n", ndev, WORLD); return 77; }
    CHECK(stb_set_device(rank));
    char id[STB_COMM_ID_BYTES];
    if (rank == 0) {
        CHECK(stb_comm_unique_id(id));
        if (write(fd_out, id, sizeof id) != (ssize_t)sizeof id) return 2;
    } else if (read(fd_in, id, sizeof id) != (ssize_t)sizeof id) return 2;
    stb_comm_t comm;
    CHECK(stb_comm_init(id, WORLD, rank, &comm));

    /* the same field and the same global launch set on both ranks */
    static double ax[(NY - 1) * (NX - 1)], ay[(NY - 1) * (NX - 1)], g[4][N];
    unsigned s = 12345u;
    for (int i = 0; i < (NY - 1) * (NX - 1); ++i) { ax[i] = 0.02 * (frand(&s) - 0.5); ay[i] = 0.02 * (frand(&s) - 0.5); }
    for (int r = 0; r < N; ++r) {
        g[0][r] = frand(&s) * (NX - 1); g[1][r] = frand(&s) * (NY - 1);
        double a = 6.2831853 * frand(&s), k = 0.3 + frand(&s);
        g[2][r] = k * cos(a); g[3][r] = k * sin(a);
    }
    stb_field_t f;
    CHECK(stb_field_create(ax, ay, NX, NY, 0, &f));

    /* my block-interleaved shard: global blocks rank, rank + WORLD, ... */
    const int64_t lblocks = stb_shard_local_blocks(N, WORLD, B), nloc = stb_shard_local_rays(N, rank, WORLD, B);
    const int64_t npad = lblocks * B;
    double *h_loc = calloc((size_t)4 * nloc, sizeof(double));
    int64_t m = 0;
    for (int64_t blk = rank; blk * B < N; blk += WORLD)
        for (int j = 0; j < B && blk * B + j < N; ++j, ++m)
            for (int c = 0; c < 4; ++c) h_loc[c * nloc + m] = g[c][blk * B + j];
    if (m != nloc) { fprintf(stderr, "rank %d: shard arithmetic %lld != %lld\n", rank, (long long)m, (long long)nloc); return 2; }

    int32_t *d_ns; double *d_end, *d_end_pad, *d_hist;
    CUCHECK(cudaMalloc((void **)&d_ns, (size_t)npad * sizeof(int32_t)));
    CUCHECK(cudaMemset(d_ns, 0, (size_t)npad * sizeof(int32_t)));
    CUCHECK(cudaMalloc((void **)&d_end, (size_t)4 * nloc * sizeof(double)));
    CUCHECK(cudaMalloc((void **)&d_end_pad, (size_t)4 * npad * sizeof(double)));
    CUCHECK(cudaMemset(d_end_pad, 0, (size_t)4 * npad * sizeof(double)));
    CUCHECK(cudaMalloc((void **)&d_hist, NBINS * sizeof(double)));
    CUCHECK(cudaMemset(d_hist, 0, NBINS * sizeof(double)));
    if (trace(rank, f, h_loc, nloc, d_ns, d_end, d_hist)) return 2;
    for (int c = 0; c < 4; ++c)     /* [4][nloc] -> [4][npad] */
        CUCHECK(cudaMemcpy(d_end_pad + (size_t)c * npad, d_end + (size_t)c * nloc, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice));

    /* exchange: histogram all-reduce; gather on rank 0; global order */
    CHECK(stb_comm_allreduce_f64(comm, d_hist, NBINS, NULL));
    int32_t *r_ns = NULL, *g_ns = NULL; double *r_end = NULL, *g_end = NULL;
    if (rank == 0) {
        CUCHECK(cudaMalloc((void **)&r_ns, (size_t)WORLD * npad * sizeof(int32_t)));
        CUCHECK(cudaMalloc((void **)&r_end, (size_t)WORLD * 4 * npad * sizeof(double)));
        CUCHECK(cudaMalloc((void **)&g_ns, (size_t)N * sizeof(int32_t)));
        CUCHECK(cudaMalloc((void **)&g_end, (size_t)4 * N * sizeof(double)));
    }
    CHECK(stb_comm_gather(comm, 0, d_ns, npad * (int64_t)sizeof(int32_t), r_ns, NULL));
    CHECK(stb_comm_gather(comm, 0, d_end_pad, 4 * npad * (int64_t)sizeof(double), r_end, NULL));
    int bad = 0;
    if (rank == 0) {
        CHECK(stb_unshuffle_blocks(r_ns, g_ns, WORLD, 1, lblocks, B, 4, 0, N, N, NULL));
        CHECK(stb_unshuffle_blocks(r_end, g_end, WORLD, 4, lblocks, B, 8, 0, N, N, NULL));
        CUCHECK(cudaDeviceSynchronize());
        /* the whole beam on one GPU */
        int32_t *w_ns; double *w_end, *w_hist;
        CUCHECK(cudaMalloc((void **)&w_ns, N * sizeof(int32_t)));
        CUCHECK(cudaMalloc((void **)&w_end, 4 * N * sizeof(double)));
        CUCHECK(cudaMalloc((void **)&w_hist, NBINS * sizeof(double)));
        CUCHECK(cudaMemset(w_hist, 0, NBINS * sizeof(double)));
        if (trace(rank, f, &g[0][0], N, w_ns, w_end, w_hist)) return 2;
        static int32_t a_ns[N], b_ns[N]; static double a_end[4 * N], b_end[4 * N]; double a_h[NBINS], b_h[NBINS];
        CUCHECK(cudaMemcpy(a_ns, g_ns, sizeof a_ns, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(b_ns, w_ns, sizeof b_ns, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(a_end, g_end, sizeof a_end, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(b_end, w_end, sizeof b_end, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(a_h, d_hist, sizeof a_h, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(b_h, w_hist, sizeof b_h, cudaMemcpyDeviceToHost));
        bad |= memcmp(a_ns, b_ns, sizeof a_ns) != 0;
        bad |= memcmp(a_end, b_end, sizeof a_end) != 0;
        double num = 0, den = 0; long long steps = 0;
        for (int i = 0; i < NBINS; ++i) { num += fabs(a_h[i] - b_h[i]); den += fabs(b_h[i]); }
        for (int i = 0; i < N; ++i) steps += b_ns[i];
        bad |= !(num <= 1e-12 * den) || den == 0;
        printf("comm_2rank: %d rays, %lld ray-steps, step counts + end states %s, histogram rel L1 %.2e -> %s\n", N, steps,
               memcmp(a_ns, b_ns, sizeof a_ns) || memcmp(a_end, b_end, sizeof a_end) ? "DIFFER" : "bit-identical",
               den ? num / den : -1.0, bad ? "MISMATCH" : "ok");
    }
    int r2 = -1, n2 = -1;
    CHECK(stb_comm_rank(comm, &r2, &n2));
    bad |= r2 != rank || n2 != WORLD;
    CHECK(stb_comm_destroy(comm));
    CHECK(stb_field_destroy(f));
    return bad ? 1 : 0;
}

int main(void)
{
    int p[2];
    if (pipe(p)) return 2;
    /* fork BEFORE any CUDA call: a forked child cannot use a CUDA runtime its parent initialised */
    pid_t pid = fork();
    if (pid < 0) return 2;
    if (pid == 0) { close(p[1]); _exit(run(1, p[0], -1)); }
    close(p[0]);
    int rc0 = run(0, -1, p[1]);
    int status = 0;
    waitpid(pid, &status, 0);
    int rc1 = WIFEXITED(status) ? WEXITSTATUS(status) : 2;
    if (rc0 || rc1) fprintf(stderr, "comm_2rank: rank 0 -> %d, rank 1 -> %d\n", rc0, rc1);
    return rc0 ? rc0 : rc1;
}
