/*
 * b200ddp.h — C ABI of libb200ddp.so: the B200-native gradient-bucket allreduce path.
 *
 * This is the drop-in boundary for the ONE hot path of subramen/distributed-pytorch: the per-step
 * gradient-bucket allreduce that `DDP(model, device_ids=[gpu_id])` (ref:multigpu.py:36,
 * ref:multigpu_torchrun.py:34, ref:multigpu_profile.py:41) fires inside `loss.backward()`
 * (ref:multigpu.py:42).  The arithmetic of that path lives in the reference's only dependency, torch
 * (ref:requirements.txt:1): c10d::Reducer + ProcessGroupNCCL.  Each entry point below names the piece
 * of that stack it replaces ("torch:" = site-packages/torch/, headers under
 * torch:include/torch/csrc/distributed/c10d/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.  Loaded with ctypes
 *     (distributed-pytorch_b200/_lib.py).
 *   - Every function returns 0 on success or a negative B200DDP_E* code; the message is available from
 *     b200ddp_last_error().  Nothing throws or aborts across the ABI.
 *   - Streams are caller-owned cudaStream_t passed as void*.  All hot-path calls only ENQUEUE work and
 *     return; they never block the host.
 *   - Host arrays (pointer/offset/length tables) only need to live for the duration of the call: they
 *     are copied into the kernel-parameter space of the launches the call makes.
 *   - One caller at a time per context (an internal mutex serialises callers; the autograd engine
 *     thread is the usual caller, like Reducer::autograd_hook under Reducer::mutex_, reducer.hpp:211).
 *   - All ranks must issue the same sequence of collective calls (allreduce / fused_bucket / barrier)
 *     with the same sizes, grid limits and algorithms: there is no tag matching, exactly like the strict
 *     bucket order of Reducer::mark_bucket_ready (reducer.hpp:231,282).
 */
#ifndef B200DDP_H_
#define B200DDP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DDP_ABI_VERSION 2

/* maximum ranks of one NVSwitch domain this library addresses (one 8-GPU box, north_star) */
#define B200DDP_MAX_WORLD 8
/* size in bytes of the opaque bootstrap handle each rank exports (b200ddp_export_handle) */
#define B200DDP_HANDLE_BYTES 128

/* error codes */
#define B200DDP_OK 0
#define B200DDP_EINVAL (-1)   /* bad argument (alignment, range, dtype, world size ...) */
#define B200DDP_ECUDA (-2)    /* a CUDA runtime / driver call failed */
#define B200DDP_ESTATE (-3)   /* call made in the wrong state (e.g. allreduce before open_peers) */
#define B200DDP_ETIMEOUT (-4) /* a peer did not arrive at an in-kernel barrier within the timeout */
#define B200DDP_EUNSUP (-5)   /* requested transport / algorithm not available on this system */
#define B200DDP_ESYS (-6)     /* an OS call (socket, sendmsg ...) failed during bootstrap */

/* element types of a bucket (buckets are per-dtype, like compute_bucket_assignment_by_size's BucketKey) */
#define B200DDP_F32 0
#define B200DDP_BF16 1
#define B200DDP_F16 2

/* how the shared arena is allocated / exchanged between the per-GPU processes */
#define B200DDP_MEM_AUTO 0 /* VMM, falling back to cudaIpc if fd export is unsupported */
#define B200DDP_MEM_IPC 1  /* cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle */
#define B200DDP_MEM_VMM 2  /* cuMemCreate + POSIX fd over a unix socket; enables NVLS multicast */

/* allreduce algorithms */
#define B200DDP_ALGO_AUTO 0
#define B200DDP_ALGO_ONESHOT 1     /* every rank reads every peer's bucket; out-of-place            */
#define B200DDP_ALGO_TWOSHOT 2     /* owner-reduces 1/ws slice by peer LDG.128, pushes result to all */
#define B200DDP_ALGO_TWOSHOT_TMA 3 /* same, peer slices staged into smem by cp.async.bulk (TMA)      */
#define B200DDP_ALGO_NVLS 4        /* multimem.ld_reduce / multimem.st through the NVSwitch          */

typedef struct b200ddp_ctx b200ddp_ctx;

/* ---- bootstrap: replaces the communicator setup behind init_process_group(backend="nccl")
 *      (ref:multigpu.py:12-20, ref:multigpu_torchrun.py:12-13 -> torch ProcessGroupNCCL) -------------- */

/* ABI version of the loaded library. */
int b200ddp_abi_version(void);

/* Create a context on CUDA device `device` for `rank` of `world` (1..B200DDP_MAX_WORLD).  Allocates the
 * shareable arena (`arena_bytes` of bucket storage, rounded up to the allocation granularity, zero
 * filled) plus the signal pads.  `mem_mode` is one of B200DDP_MEM_*. */
int b200ddp_init(int rank, int world, int device, uint64_t arena_bytes, int mem_mode, b200ddp_ctx** out);

/* Write this rank's B200DDP_HANDLE_BYTES bootstrap handle to `out`.  The caller moves the handles
 * between processes (the c10d store in our Python host) and gives every rank the concatenation. */
int b200ddp_export_handle(b200ddp_ctx* ctx, void* out, uint64_t cap);

/* Collective over all ranks: map every peer's arena (`all_handles` = world * B200DDP_HANDLE_BYTES,
 * ordered by rank) and, in VMM mode when the fabric supports it, build the NVLS multicast mapping.
 * Handles carrying this process's pid are opened as plain pointers (single-process loop-back worlds
 * used by the 1-GPU parity tests). */
int b200ddp_open_peers(b200ddp_ctx* ctx, const void* all_handles, int n);

/* Device pointer / size of this rank's bucket arena (offsets in the calls below are relative to it). */
int b200ddp_arena_ptr(b200ddp_ctx* ctx, void** dev_ptr);
uint64_t b200ddp_arena_bytes(b200ddp_ctx* ctx);
/* 1 if B200DDP_ALGO_NVLS is usable (multicast mapping was built), else 0. */
int b200ddp_has_multicast(b200ddp_ctx* ctx);
/* B200DDP_MEM_IPC or B200DDP_MEM_VMM: the transport actually in use. */
int b200ddp_mem_mode(b200ddp_ctx* ctx);

/* Unmap peers, free the arena.  Callers fence with a host barrier first (destroy_process_group path,
 * ref:multigpu.py:88). */
int b200ddp_destroy(b200ddp_ctx* ctx);

/* Last error message of `ctx` (or of the calling thread's last failed init when ctx == NULL). */
const char* b200ddp_last_error(b200ddp_ctx* ctx);

/* ---- hot path ------------------------------------------------------------------------------------ */

/* Flatten + pre-scale one bucket: for each of the `n` tensors, arena[bucket_off + off_el[i] ...] <-
 * grad_i * scale, element type `dtype`.  Replaces Reducer::mark_variable_ready_dense (reducer.hpp:276;
 * the per-parameter `bucket_view_in.copy_(grad) * 1/ws` D2D launches) with ONE multi-tensor launch per
 * <=B200DDP table chunk.  `bucket_off` is in bytes and 16-byte aligned; `off_el`/`len_el` are in
 * elements; tensors must be contiguous and must not overlap inside the bucket. */
int b200ddp_flatten_scale(b200ddp_ctx* ctx, const void* const* grad_ptrs, const uint64_t* off_el,
                          const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, float scale,
                          void* stream);

/* Unflatten: grad_i <- arena[bucket_off + off_el[i] ...].  Replaces Reducer::finalize_bucket_dense ->
 * copy_bucket_to_grad (reducer.hpp:284,330). */
int b200ddp_unflatten(b200ddp_ctx* ctx, void* const* grad_ptrs, const uint64_t* off_el,
                      const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, void* stream);

/* SUM-allreduce `nbytes` (multiple of 16) of arena[in_off] over all ranks into arena[out_off] on every
 * rank (in_off == out_off allowed except for ONESHOT), accumulating in fp32; the reduced sum is multiplied by
 * `scale` (fp32, one rounding; 1.0 = none) before it is written back.  The Reducer passes 1.0 because it
 * pre-scales in flatten_scale exactly like the reference (reducer.hpp:276 divides before the allreduce); the
 * comm-hook shim passes 1/world, since torch hands a hook an UNdivided bucket
 * (torch:nn/parallel/distributed.py:2021-2023).  Replaces
 * Reducer::all_reduce_bucket -> _AllReduceBySumCommHook -> ProcessGroupNCCL::allreduce -> ncclAllReduce
 * (reducer.hpp:524, default_comm_hooks.hpp:41-51, ProcessGroupNCCL.hpp:856).  The kernel itself issues
 * the NVLink traffic (peer loads/stores, or multimem for NVLS); `max_ctas` caps the grid (0 = library
 * default) so the backward's GEMMs keep the remaining SMs.  With world == 1 this is a no-op copy when
 * in_off != out_off. */
int b200ddp_allreduce(b200ddp_ctx* ctx, uint64_t in_off, uint64_t out_off, uint64_t nbytes, int dtype,
                      int algo, int max_ctas, float scale, void* stream);

/* Small-bucket fast path: flatten+scale, one-shot allreduce and unflatten of one bucket in ONE launch
 * (n <= b200ddp_max_table_tensors(), bucket bytes <= b200ddp_fused_max_bytes()).  Equivalent to
 * flatten_scale -> allreduce(ONESHOT) -> unflatten.  With world > 1 the arena staging is double buffered so
 * that the kernel needs a single cross-GPU barrier, and buckets of at most 32 KiB use a barrier-free LL
 * protocol (4-byte word + 4-byte tag per 8-byte NVLink store): the bucket then occupies
 * [bucket_off, bucket_off + b200ddp_fused_region_bytes(ctx, bucket_bytes)), which the caller must have
 * cleared once with b200ddp_clear() before the first use (and after re-using the bytes for anything else). */
int b200ddp_fused_bucket(b200ddp_ctx* ctx, void* const* grad_ptrs, const uint64_t* off_el,
                         const uint64_t* len_el, int n, int dtype, uint64_t bucket_off,
                         uint64_t bucket_bytes, float scale, void* stream);
int b200ddp_max_table_tensors(void);
uint64_t b200ddp_fused_max_bytes(void);

/* Optional optimizer epilogue (opt-in from the host: DistributedDataParallel.fuse_optimizer).  The same
 * two calls with one more pointer table: wherever the kernel writes the averaged gradient g_i back it also applies plain
 * SGD to the parameter the gradient belongs to, param_i <- fma(-lr, g_i, param_i) — what torch.optim.SGD(lr) without
 * momentum / weight decay computes in its own pass over all parameters (ref:multigpu.py:43,68).  param_ptrs[i] may be
 * NULL (that tensor is only un-flattened); parameters are laid out like their gradients (same dtype, same length). */
int b200ddp_unflatten_sgd(b200ddp_ctx* ctx, void* const* grad_ptrs, void* const* param_ptrs, const uint64_t* off_el,
                          const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, float lr, void* stream);
int b200ddp_fused_bucket_sgd(b200ddp_ctx* ctx, void* const* grad_ptrs, void* const* param_ptrs,
                             const uint64_t* off_el, const uint64_t* len_el, int n, int dtype, uint64_t bucket_off,
                             uint64_t bucket_bytes, float scale, float lr, void* stream);
uint64_t b200ddp_fused_region_bytes(b200ddp_ctx* ctx, uint64_t bucket_bytes);
/* the same for a world size given explicitly (no context: callers size the arena BEFORE b200ddp_init) */
uint64_t b200ddp_fused_region_bytes_for(int world, uint64_t bucket_bytes);

/* Diagnostic: arena[dst_off] of this rank <- arena[src_off] of rank `peer`, `nbytes` (multiple of 16) pulled
 * with 16-byte loads over the peer mapping.  No synchronisation with the peer: it measures the per-direction
 * NVLink read bandwidth a kernel can reach on this box (the denominator bench.py reports as
 * `nvlink_peak_measured`); `peer` == own rank measures the local HBM copy. */
int b200ddp_peer_read(b200ddp_ctx* ctx, int peer, uint64_t src_off, uint64_t dst_off, uint64_t nbytes,
                      int max_ctas, void* stream);

/* Launch an instantiated CUDA graph (a captured training step whose nodes include this library's kernels) on
 * `stream` and, when `done_event` is not NULL, record it right behind: the two driver calls a replayed step costs,
 * without the interpreter between them.  Handles are the raw cudaGraphExec_t / cudaStream_t / cudaEvent_t values.
 * Context-free; errors via b200ddp_last_error(NULL). */
int b200ddp_launch_graph(void* graph_exec, void* stream, void* done_event);

/* Input pipeline (replaces the per-batch `source.to(gpu_id)` of ref:multigpu.py:48-50 for a shard that
 * is resident in HBM): dst[b] = src[(*cursor + b) mod full_rows] for b < batch_rows, rows of `row_bytes` (a
 * multiple of 4), then *cursor = (*cursor + batch_rows) mod full_rows.  With `split_bytes` > 0 the first split_bytes
 * of every row land in `dst` as a dense [batch_rows, split_bytes] block and the rest in `dst2` (source and targets of a
 * packed [source | targets] shard come out as two contiguous tensors).  `cursor` is a device uint64; no host state, so
 * a captured training step replays it.  Context-free (no arena involved); errors via b200ddp_last_error(NULL). */
int b200ddp_gather_rows(const void* src, void* dst, void* dst2, uint64_t row_bytes, uint64_t split_bytes,
                        uint32_t batch_rows, uint64_t full_rows, void* cursor, void* stream);

/* Zero arena[off, off + nbytes) on `stream` (offset and size multiples of 16). */
int b200ddp_clear(b200ddp_ctx* ctx, uint64_t off, uint64_t nbytes, void* stream);

/* Broadcast `nbytes` (multiple of 16) of arena[off] from rank `root` to arena[off] of every rank by
 * peer stores.  Used for DDP's construction-time parameter sync and per-forward buffer sync
 * (torch:nn/parallel/distributed.py:881-890,2214-2256). */
int b200ddp_broadcast(b200ddp_ctx* ctx, uint64_t off, uint64_t nbytes, int root, int max_ctas,
                      void* stream);

/* In-kernel barrier over all ranks on `stream` (signal pads, release/acquire at .sys scope). */
int b200ddp_barrier(b200ddp_ctx* ctx, void* stream);

/* Make `waiter` wait (on the device) for everything enqueued so far on `signaled`: the event
 * record/wait pair the Reducer uses to order its side stream against the autograd stream. */
int b200ddp_stream_wait(b200ddp_ctx* ctx, void* waiter, void* signaled);

/* ---- diagnostics --------------------------------------------------------------------------------- */

/* Returns 0, or B200DDP_ETIMEOUT if any kernel since the last poll observed a peer timeout (the flag
 * lives in mapped host memory; polling does not synchronise the device).  A timeout is sticky on the device
 * side: the rank's later collective kernels return at once and unflatten launches leave the gradients
 * untouched, so a half-reduced bucket is never scattered; the context must be destroyed afterwards. */
int b200ddp_poll_error(b200ddp_ctx* ctx);
/* In-kernel barrier timeout in milliseconds (default 600000 = 10 minutes, the order of NCCL's watchdog: a peer
 * may legitimately arrive late, e.g. while rank 0 writes a checkpoint). */
int b200ddp_set_timeout_ms(b200ddp_ctx* ctx, uint32_t ms);
/* Number of kernels this context has launched since init. */
uint64_t b200ddp_launch_count(b200ddp_ctx* ctx);
/* The algorithm AUTO resolves to for a bucket of `nbytes` on this context. */
int b200ddp_pick_algo(b200ddp_ctx* ctx, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif /* B200DDP_H_ */
