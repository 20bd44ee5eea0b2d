// plan_b200.cu -- exactly-sized, length-sorted trajectory arenas (stb_traj_plan, include/sidtrace_b200.h).
//
// The reference hands every ray four MaxStep-long arrays (ray.py:19-22) and keeps [:i] (ray.py:99-102).
// A batch that wants every kept state in HBM cannot afford rays x cap states when step counts are
// heavy-tailed (a trapped ray runs 20x the median), and a tile of rays in lock-step rows idles while its
// longest ray finishes.  With the step counts of a first, O(rays)-output pass the layout is planned
// exactly: rays sorted by step count (longest first, ties in launch order), T consecutive ones per tile,
// a tile owns (its longest ray's states) rows of T x 32 B.  Linked into libsidtrace_b200.so.
#include "../../include/sidtrace_b200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdint>
#include <cstdio>
#include <mutex>

extern "C" int stb_internal_fail(int code, const char *msg);
extern "C" void stb_internal_count_launch(void);

namespace {

int failf(int code, const char *what, cudaError_t e)
{
    char buf[256];
    snprintf(buf, sizeof(buf), "stb_traj_plan: %s: %s", what, cudaGetErrorString(e));
    return stb_internal_fail(code, buf);
}

#define PCU(call)                                                            \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) { rc = failf(e_ == cudaErrorMemoryAllocation ? STB_ERR_NOMEM : STB_ERR_CUDA, #call, e_); goto done; } \
    } while (0)

// Scratch for the sort comes from a memory pool of the library's own whose release threshold is "never": the
// device's default pool hands its memory back at every synchronisation, which made every plan pay 5-25 ms of
// allocation for 0.2 ms of work (profiles/r02/plan_pool.txt).
std::mutex g_pool_mu;
cudaMemPool_t g_pool[64] = {};
cudaError_t scratch_pool(cudaMemPool_t *out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool[dev]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if ((e = cudaMemPoolCreate(&g_pool[dev], &props)) != cudaSuccess) return e;
        uint64_t keep = UINT64_MAX;
        if ((e = cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
    }
    *out = g_pool[dev];
    return cudaSuccess;
}

__global__ void iota_kernel(int *v, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) v[i] = (int)i;
}

// rows of tile t = states of its longest ray = (largest step count in the tile) + 1: states 0..i are
// written (sidtracemodule.c:36-48).  Keys are sorted in descending order, so that is the first slot.
__global__ void tile_rows_kernel(const int *keys_sorted, int tile, long long ntiles, long long *rows)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; t <= ntiles; t += step) rows[t] = t < ntiles ? (long long)keys_sorted[t * tile] + 1 : 0;
}

__global__ void perm_kernel(const int *vals_sorted, long long nrays, long long nslots, int *perm, int *slot_of)
{
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; s < nslots; s += step) {
        const int r = s < nrays ? vals_sorted[s] : -1;
        perm[s] = r;
        if (r >= 0) slot_of[r] = (int)s;
    }
}

}  // namespace

extern "C" int stb_traj_plan(const int32_t *nsteps, int64_t nrays, int32_t tile, int64_t L,
                             int32_t *perm, int32_t *slot_of, int64_t *tile_off, int64_t *total_rows, void *stream)
{
    if (!nsteps || !perm || !slot_of || !tile_off || !total_rows) return stb_internal_fail(STB_ERR_ARG, "stb_traj_plan: NULL argument");
    if (tile != 32 && tile != 64 && tile != 128 && tile != 256)
        return stb_internal_fail(STB_ERR_ARG, "stb_traj_plan: tile must be 32, 64, 128 or 256");
    if (nrays < 0 || nrays > 0x7fffffffLL) return stb_internal_fail(STB_ERR_ARG, "stb_traj_plan: nrays out of range");
    *total_rows = 0;
    if (nrays == 0) return STB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const long long ntiles = (nrays + tile - 1) / tile, nslots = ntiles * tile;
    int end_bit = 32;
    if (L > 0 && L <= 0x7fffffffLL) { end_bit = 1; while ((1LL << end_bit) < L) ++end_bit; }

    int rc = STB_OK;
    int *keys_out = nullptr, *vals_in = nullptr, *vals_out = nullptr;
    long long *rows = nullptr;
    void *tmp = nullptr;
    cudaMemPool_t pool = nullptr;
    size_t tmp_sort = 0, tmp_scan = 0, tmp_bytes = 0;
    long long total = 0;
    const int threads = 256;
    const unsigned blocks = (unsigned)((nslots + threads - 1) / threads < 148 * 16 ? (nslots + threads - 1) / threads : 148 * 16);

    PCU(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_sort, nsteps, keys_out, vals_in, vals_out, (int)nrays, 0, end_bit, s));
    PCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, rows, (long long *)tile_off, (int)(ntiles + 1), s));
    tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
    PCU(scratch_pool(&pool));
    PCU(cudaMallocFromPoolAsync((void **)&keys_out, (size_t)nrays * sizeof(int), pool, s));
    PCU(cudaMallocFromPoolAsync((void **)&vals_in, (size_t)nrays * sizeof(int), pool, s));
    PCU(cudaMallocFromPoolAsync((void **)&vals_out, (size_t)nrays * sizeof(int), pool, s));
    PCU(cudaMallocFromPoolAsync((void **)&rows, (size_t)(ntiles + 1) * sizeof(long long), pool, s));
    PCU(cudaMallocFromPoolAsync(&tmp, tmp_bytes ? tmp_bytes : 16, pool, s));

    iota_kernel<<<blocks, threads, 0, s>>>(vals_in, nrays);
    stb_internal_count_launch();
    // stable LSD radix sort: equal step counts keep launch order
    PCU(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_sort, nsteps, keys_out, vals_in, vals_out, (int)nrays, 0, end_bit, s));
    tile_rows_kernel<<<blocks, threads, 0, s>>>(keys_out, tile, ntiles, rows);
    stb_internal_count_launch();
    PCU(cub::DeviceScan::ExclusiveSum(tmp, tmp_scan, rows, (long long *)tile_off, (int)(ntiles + 1), s));
    perm_kernel<<<blocks, threads, 0, s>>>(vals_out, nrays, nslots, perm, slot_of);
    stb_internal_count_launch();
    PCU(cudaGetLastError());
    PCU(cudaMemcpyAsync(&total, tile_off + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PCU(cudaStreamSynchronize(s));
    *total_rows = total;
done:
    if (keys_out) cudaFreeAsync(keys_out, s);
    if (vals_in) cudaFreeAsync(vals_in, s);
    if (vals_out) cudaFreeAsync(vals_out, s);
    if (rows) cudaFreeAsync(rows, s);
    if (tmp) cudaFreeAsync(tmp, s);
    return rc;
}
