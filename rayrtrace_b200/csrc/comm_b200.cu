// comm_b200.cu -- multi-GPU entry points of the C ABI (include/sidtrace_b200_comm.h): NCCL over
// NVLink 5 / NVSwitch, loaded at run time; block-interleaved sharding; the unshuffle kernel.
#include "../../include/sidtrace_b200_comm.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

extern "C" int stb_internal_fail(int code, const char *msg);
extern "C" void stb_internal_count_launch(void);

namespace {

struct Nccl {
    void *h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

Nccl g_nccl;
std::once_flag g_nccl_once;

void load_nccl()
{
    // the copy the process already has (torch loads its bundled libnccl.so.2), else the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h)
        for (const char *n : names) {
            g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.h) break;
        }
    if (!g_nccl.h) return;
#define SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.h, #name); if (!g_nccl.field) return;
    SYM(GetUniqueId, ncclGetUniqueId)
    SYM(CommInitRank, ncclCommInitRank)
    SYM(CommDestroy, ncclCommDestroy)
    SYM(AllReduce, ncclAllReduce)
    SYM(AllGather, ncclAllGather)
    SYM(Send, ncclSend)
    SYM(Recv, ncclRecv)
    SYM(GroupStart, ncclGroupStart)
    SYM(GroupEnd, ncclGroupEnd)
    SYM(GetErrorString, ncclGetErrorString)
#undef SYM
    g_nccl.ok = true;
}

int need_nccl()
{
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) return stb_internal_fail(STB_ERR_COMM, "libnccl.so.2 could not be loaded (dlopen): the multi-GPU entry points need NCCL");
    return STB_OK;
}

int nfail(const char *what, ncclResult_t r)
{
    char buf[256];
    snprintf(buf, sizeof(buf), "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return stb_internal_fail(STB_ERR_COMM, buf);
}

#define NC(call)                                             \
    do {                                                     \
        ncclResult_t r_ = (call);                            \
        if (r_ != ncclSuccess) return nfail(#call, r_);      \
    } while (0)

template <typename T>
__global__ void unshuffle_kernel(const T *__restrict__ src, T *__restrict__ dst, int world, int rows, long long lblocks,
                                 int block, long long g0, long long n_global, long long dst_stride)
{
    // one thread per element of src[world][rows][lblocks*block]; consecutive threads = consecutive rays of
    // one block on both sides
    const long long nloc = lblocks * block;
    const long long total = (long long)world * rows * nloc;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += step) {
        const long long j = i % nloc;
        const long long rr = i / nloc;
        const int row = (int)(rr % rows);
        const int r = (int)(rr / rows);
        const long long lb = j / block;
        const long long g = g0 + (lb * world + r) * block + (j - lb * block);
        if (g < n_global) dst[(long long)row * dst_stride + g] = src[i];
    }
}

// The same copy, one CTA per (local block, row, rank): a block of one row is a contiguous run of `block` elements on
// both sides, so it moves as 16-byte vectors with the index arithmetic done once per CTA instead of five 64-bit
// divisions per element (the element-wise kernel above spent ~6 ms of full-GPU instruction time on a 3.6 GB gather,
// on the root, under its trace).  Needs 16-byte-aligned runs; the caller falls back to the kernel above otherwise.
__global__ void unshuffle_runs_kernel(const char *__restrict__ src, char *__restrict__ dst, int world, int rows, long long lblocks,
                                      int run_bytes, int elem_bytes, long long g0, long long n_global, long long dst_stride)
{
    const long long lb = blockIdx.x;
    const int row = blockIdx.y, r = blockIdx.z;
    const long long g = g0 + (lb * world + r) * (long long)(run_bytes / elem_bytes);          // first global ray of the run
    if (g >= n_global) return;
    long long live = (n_global - g) * elem_bytes;                                              // bytes of the run that exist
    if (live > run_bytes) live = run_bytes;
    const char *s = src + (((long long)r * rows + row) * lblocks + lb) * run_bytes;
    char *d = dst + ((long long)row * dst_stride + g) * elem_bytes;
    for (int o = threadIdx.x * 16; o < run_bytes; o += blockDim.x * 16) {
        if (o + 16 <= live) {
            *reinterpret_cast<int4 *>(d + o) = *reinterpret_cast<const int4 *>(s + o);
        } else {
            for (int b = o; b < live; b += elem_bytes) {                                        // the one partial run of the launch
                if (elem_bytes == 4) *reinterpret_cast<int *>(d + b) = *reinterpret_cast<const int *>(s + b);
                else *reinterpret_cast<double *>(d + b) = *reinterpret_cast<const double *>(s + b);
            }
        }
    }
}

}  // namespace

struct stb_comm_s {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

extern "C" {

int stb_comm_unique_id(void *id)
{
    if (!id) return stb_internal_fail(STB_ERR_ARG, "stb_comm_unique_id: NULL argument");
    int rc = need_nccl();
    if (rc != STB_OK) return rc;
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    static_assert(sizeof(ncclUniqueId) == STB_COMM_ID_BYTES, "unique id size");
    return STB_OK;
}

int stb_comm_init(const void *id, int nranks, int rank, stb_comm_t *out)
{
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return stb_internal_fail(STB_ERR_ARG, "stb_comm_init: bad argument");
    int rc = need_nccl();
    if (rc != STB_OK) return rc;
    stb_comm_s *c = new (std::nothrow) stb_comm_s;
    if (!c) return stb_internal_fail(STB_ERR_NOMEM, "stb_comm_init: host allocation failed");
    c->rank = rank; c->nranks = nranks;
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; return stb_internal_fail(STB_ERR_CUDA, "stb_comm_init: no CUDA device"); }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) { delete c; return nfail("ncclCommInitRank", r); }
    *out = c;
    return STB_OK;
}

int stb_comm_destroy(stb_comm_t c)
{
    if (!c) return STB_OK;
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    delete c;
    return STB_OK;
}

int stb_comm_rank(stb_comm_t c, int *rank, int *nranks)
{
    if (!c) return stb_internal_fail(STB_ERR_ARG, "stb_comm_rank: NULL communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return STB_OK;
}

int stb_comm_allreduce_f64(stb_comm_t c, double *buf, int64_t n, void *stream)
{
    if (!c || (!buf && n > 0) || n < 0) return stb_internal_fail(STB_ERR_ARG, "stb_comm_allreduce_f64: bad argument");
    if (n == 0) return STB_OK;
    NC(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream));
    return STB_OK;
}

int stb_comm_allreduce_max_f64(stb_comm_t c, double *buf, int64_t n, void *stream)
{
    if (!c || (!buf && n > 0) || n < 0) return stb_internal_fail(STB_ERR_ARG, "stb_comm_allreduce_max_f64: bad argument");
    if (n == 0) return STB_OK;
    NC(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclMax, c->comm, (cudaStream_t)stream));
    return STB_OK;
}

int stb_comm_gather(stb_comm_t c, int root, const void *send, int64_t bytes, void *recv, void *stream)
{
    if (!c || root < 0 || root >= c->nranks || bytes < 0 || (bytes > 0 && !send) || (bytes > 0 && c->rank == root && !recv))
        return stb_internal_fail(STB_ERR_ARG, "stb_comm_gather: bad argument");
    if (bytes == 0) return STB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // NCCL has no gather: one grouped set of sends and receives (a single fused launch)
    NC(g_nccl.GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->nranks; ++r) {
            char *dst = (char *)recv + (size_t)r * (size_t)bytes;
            if (r == root) {
                if (cudaMemcpyAsync(dst, send, (size_t)bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
                    g_nccl.GroupEnd();
                    return stb_internal_fail(STB_ERR_CUDA, "stb_comm_gather: local copy failed");
                }
            } else {
                ncclResult_t r_ = g_nccl.Recv(dst, (size_t)bytes, ncclChar, r, c->comm, s);
                if (r_ != ncclSuccess) { g_nccl.GroupEnd(); return nfail("ncclRecv", r_); }
            }
        }
    } else {
        ncclResult_t r_ = g_nccl.Send(send, (size_t)bytes, ncclChar, root, c->comm, s);
        if (r_ != ncclSuccess) { g_nccl.GroupEnd(); return nfail("ncclSend", r_); }
    }
    NC(g_nccl.GroupEnd());
    return STB_OK;
}

int stb_comm_allgather(stb_comm_t c, const void *send, int64_t bytes, void *recv, void *stream)
{
    if (!c || bytes < 0 || (bytes > 0 && (!send || !recv))) return stb_internal_fail(STB_ERR_ARG, "stb_comm_allgather: bad argument");
    if (bytes == 0) return STB_OK;
    NC(g_nccl.AllGather(send, recv, (size_t)bytes, ncclChar, c->comm, (cudaStream_t)stream));
    return STB_OK;
}

int64_t stb_shard_local_blocks(int64_t n_global, int32_t world, int32_t block)
{
    if (n_global <= 0 || world < 1 || block < 1) return 0;
    const int64_t nb = (n_global + block - 1) / block;
    return (nb + world - 1) / world;
}

int64_t stb_shard_local_rays(int64_t n_global, int32_t rank, int32_t world, int32_t block)
{
    if (n_global <= 0 || world < 1 || block < 1 || rank < 0 || rank >= world) return 0;
    const int64_t nb = (n_global + block - 1) / block;          // global blocks, the last one maybe partial
    int64_t mine = nb / world + (rank < nb % world ? 1 : 0);    // blocks rank, rank + world, ...
    if (mine == 0) return 0;
    int64_t rays = mine * block;
    const int64_t last = nb - 1;
    if (last % world == rank) rays -= nb * (int64_t)block - n_global;   // the partial block is mine
    return rays;
}

int stb_unshuffle_blocks(const void *src, void *dst, int32_t world, int32_t rows, int64_t lblocks, int32_t block,
                         int32_t elem_bytes, int64_t g0, int64_t n_global, int64_t dst_stride, void *stream)
{
    if (!src || !dst || world < 1 || rows < 1 || lblocks < 0 || block < 1 || (elem_bytes != 4 && elem_bytes != 8) || g0 < 0 ||
        dst_stride < n_global)
        return stb_internal_fail(STB_ERR_ARG, "stb_unshuffle_blocks: bad argument");
    const long long total = (long long)world * rows * lblocks * block;
    if (total == 0) return STB_OK;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return stb_internal_fail(STB_ERR_CUDA, "stb_unshuffle_blocks: no CUDA device");
    // whole runs as 16-byte vectors when every run starts 16-byte aligned on both sides
    const long long run_bytes = (long long)block * elem_bytes;
    const bool aligned = run_bytes % 16 == 0 && run_bytes <= (1 << 20) && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0 &&
                         (dst_stride * elem_bytes) % 16 == 0 && (g0 * elem_bytes) % 16 == 0 && lblocks <= 0x7fffffffLL && rows <= 65535 &&
                         world <= 65535;
    if (aligned) {
        int threads = (int)(run_bytes / 16);
        threads = threads < 32 ? 32 : (threads > 256 ? 256 : threads);
        const dim3 grid((unsigned)lblocks, (unsigned)rows, (unsigned)world);
        unshuffle_runs_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>((const char *)src, (char *)dst, world, rows, lblocks, (int)run_bytes,
                                                                            elem_bytes, g0, n_global, dst_stride);
        stb_internal_count_launch();
        if (cudaGetLastError() != cudaSuccess) return stb_internal_fail(STB_ERR_CUDA, "stb_unshuffle_blocks: launch failed");
        return STB_OK;
    }
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    if (elem_bytes == 4)
        unshuffle_kernel<int><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int *)src, (int *)dst, world, rows, lblocks, block, g0, n_global, dst_stride);
    else
        unshuffle_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const double *)src, (double *)dst, world, rows, lblocks, block, g0, n_global, dst_stride);
    stb_internal_count_launch();
    if (cudaGetLastError() != cudaSuccess) return stb_internal_fail(STB_ERR_CUDA, "stb_unshuffle_blocks: launch failed");
    return STB_OK;
}

}  // extern "C"
