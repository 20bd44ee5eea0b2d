/*
 * sidtrace_b200_comm.h -- multi-GPU entry points of the C ABI (same library as sidtrace_b200.h).
 *
 * The reference is a serial loop over rays (beam.py:340-342) with no communication of any kind; its
 * beam-level consumers (Beam.field_r, beam.py:347-444; field_result, beam.py:447-472) need every ray's
 * O(1) summary in one place.  Here a beam is sharded by ray over one process per GPU, the field is
 * replicated, and exactly two exchange steps exist:
 *   - the sum of the per-rank deposition histograms (extension)        -> stb_comm_allreduce_f64
 *   - the per-ray summaries (step counts, end states, ...) of all shards, in global ray order, on one
 *     rank or on all of them                                             -> stb_comm_gather / _allgather
 *                                                                            + stb_unshuffle_blocks
 * Trajectories are never exchanged (10^8 rays x 10^3 steps x 32 B).
 *
 * Transport: NCCL (NVLink 5 / NVSwitch), loaded at run time with dlopen("libnccl.so.2") -- the copy the
 * process already has (torch's bundled one under Python) or the system library for a plain-C host.  No
 * torch types: communicators are created from a 128-byte unique id that the host program distributes
 * however it likes (torch.distributed's store under Python; a pipe in tests/c/comm_2rank.c).
 *
 * Sharding convention ("block-interleaved", stb_shard_*): the global ray index space is cut into blocks of
 * B rays (default 256); block b belongs to rank b % world and is that rank's local block b / world.
 * Over-dense launch angles, which cluster in ray index, spread over all GPUs, and putting the gathered
 * shards back in global order is a copy of whole B-ray runs (stb_unshuffle_blocks), not a scatter.
 * Every rank's shard is padded to the same number of local blocks (the padding is never traced).
 */
#ifndef SIDTRACE_B200_COMM_H
#define SIDTRACE_B200_COMM_H

#include "sidtrace_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define STB_COMM_ID_BYTES 128
#define STB_ERR_COMM  -5      /* NCCL missing or an NCCL call failed (message in stb_last_error) */

typedef struct stb_comm_s *stb_comm_t;

/* rank 0: fill id[128] (ncclGetUniqueId); distribute it to the other ranks out of band */
int stb_comm_unique_id(void *id);
/* every rank, on its own current device (stb_set_device first): collective, blocks until all joined */
int stb_comm_init(const void *id, int nranks, int rank, stb_comm_t *out);
int stb_comm_destroy(stb_comm_t c);
int stb_comm_rank(stb_comm_t c, int *rank, int *nranks);

/* in-place sum over ranks of buf[n] (DEVICE, float64): the deposition histogram */
int stb_comm_allreduce_f64(stb_comm_t c, double *buf, int64_t n, void *stream);
/* `bytes` bytes from every rank (DEVICE) to recv[nranks][bytes] on `root` (recv ignored elsewhere) */
int stb_comm_gather(stb_comm_t c, int root, const void *send, int64_t bytes, void *recv, void *stream);
/* the same to every rank */
int stb_comm_allgather(stb_comm_t c, const void *send, int64_t bytes, void *recv, void *stream);
/* max over ranks of one float64 (DEVICE): device-timed "max over ranks" without the host */
int stb_comm_allreduce_max_f64(stb_comm_t c, double *buf, int64_t n, void *stream);

/* ---- block-interleaved sharding (host arithmetic, no device work) ---- */
/* local blocks every rank holds after padding: ceil(ceil(n_global/B) / world) */
int64_t stb_shard_local_blocks(int64_t n_global, int32_t world, int32_t block);
/* rays of `rank`'s shard that exist (the rest of its padded shard is not traced) */
int64_t stb_shard_local_rays(int64_t n_global, int32_t rank, int32_t world, int32_t block);

/* Gathered shards -> global ray order, one launch.  src[world][rows][nloc] (what stb_comm_gather left
 * on the root for a per-ray array of `rows` rows, e.g. rows = 4 for end states, nloc = lblocks * block
 * padded local rays of ONE chunk) is copied to dst[rows][dst_stride] starting at global ray `g0` =
 * first global ray of the chunk: local block lb of rank r lands at g0 + (lb * world + r) * block.
 * Elements are elem_bytes = 4 or 8 wide; rays >= n_global are dropped.  DEVICE pointers. */
int stb_unshuffle_blocks(const void *src, void *dst, int32_t world, int32_t rows, int64_t lblocks, int32_t block,
                         int32_t elem_bytes, int64_t g0, int64_t n_global, int64_t dst_stride, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIDTRACE_B200_COMM_H */
