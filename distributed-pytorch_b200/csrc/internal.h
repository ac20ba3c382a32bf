// internal.h — types shared by the host side (host.cu) and the kernels (kernels.cu) of libb200ddp.so.
// Nothing here is part of the C ABI (include/b200ddp.h is).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/b200ddp.h"

namespace b200ddp {

constexpr int kMaxWorld = B200DDP_MAX_WORLD;
// One flag slot per (CTA index, source rank); grids of collective kernels are capped to this many CTAs.
constexpr int kMaxBlocks = 512;
// Tensors one multi-tensor launch can address (its table travels in the kernel-parameter space, so a
// launch needs no device-side table memory and no H2D copy; see kernels.cu: TensorTable).
constexpr int kMaxTableTensors = 160;
// Largest bucket the single-launch fused path takes.
constexpr uint64_t kFusedMaxBytes = 256u << 10;
// Buckets up to this size use the barrier-free LL protocol inside the fused path (world > 1).
constexpr uint64_t kLLMaxBytes = 32u << 10;
// Arena bytes the fused path needs for a bucket of `bucket_bytes` (multiple of 16) in a world of `ws`:
// ws == 1: the bucket itself; LL: 2 slot sets x ws sources x (bucket_bytes / 4) words x 8 bytes;
// otherwise two staging slots.
constexpr uint64_t fused_region_bytes(uint64_t bucket_bytes, int ws) {
  return ws <= 1 ? bucket_bytes : (bucket_bytes <= kLLMaxBytes ? 4ull * ws * bucket_bytes : 2ull * bucket_bytes);
}

// Layout of the per-rank signal region that precedes the bucket arena inside the shared allocation.
//   [0, kFlagBytes)              uint32 flag[kMaxBlocks][kMaxWorld]   written by peers, polled locally
//   [kFlagBytes, +kEpochBytes)   uint32 epoch[kMaxBlocks]             local, per-CTA barrier generation
constexpr uint64_t kFlagBytes = uint64_t(kMaxBlocks) * kMaxWorld * 4;
constexpr uint64_t kEpochBytes = uint64_t(kMaxBlocks) * 4;
//   [kFlagBytes + kEpochBytes]   uint32 poisoned                      local, sticky: a kernel timed out on a peer
constexpr uint64_t kSignalBytes = 2u << 20;  // one 2 MiB page: keeps the arena 2 MiB aligned
static_assert(kFlagBytes + kEpochBytes + 4 <= kSignalBytes, "signal region too small");

// Base addresses of every rank's shared allocation as mapped into THIS process.
struct PeerTable {
  uint64_t base[kMaxWorld];  // signal region at base, arena at base + kSignalBytes
};

// Per-launch control block.
struct Ctl {
  uint32_t* err;        // mapped host word: set to non-zero by a kernel that timed out on a peer
  uint64_t timeout_ns;  // in-kernel barrier timeout
  int rank;
  int world;
};

// Multi-tensor table, passed BY VALUE as a kernel parameter (< 4 KiB with the other arguments; the opt-in SGD
// instantiations add an SgdTable and use the > 4 KiB parameter space CUDA 12.1+ offers on sm_70 and later).
// Tensor i occupies bucket vectors [off_vec[i], off_vec[i+1]) (16-byte units from the bucket start) and
// holds len_el[i] elements starting at ptr[i].  off_vec is ascending; off_vec[n] closes the last tensor.
struct TensorTable {
  uint64_t ptr[kMaxTableTensors];
  uint64_t len_el[kMaxTableTensors];
  uint32_t off_vec[kMaxTableTensors + 1];
  int32_t n;
};

// Optional optimizer epilogue (opt-in): parameter i of the bucket's tensor table lives at ptr[i];
// wherever a kernel writes the averaged gradient g back it also applies plain SGD, p <- fma(-lr, g, p) (what
// torch.optim.SGD's `param.add_(grad, alpha=-lr)` computes, ref:multigpu.py:43,68).  Travels in the kernel-parameter
// space like the tensor table; kernels are instantiated with and without it, so the default launches carry none of it.
struct SgdTable {
  uint64_t ptr[kMaxTableTensors];
  float neg_lr;
};

// kernels.cu — launch wrappers.  All return cudaError_t from the launch.  `sgd` == nullptr: no optimizer epilogue.
// `poison`: the rank's sticky timeout word (device pointer) — an unflatten launch returns without writing when
// it is set; nullptr = no check
cudaError_t launch_flatten(const TensorTable& tt, void* bucket, uint32_t vec_begin, uint32_t vec_end, int dtype,
                           float scale, bool unflatten, const uint32_t* poison, cudaStream_t s,
                           const SgdTable* sgd = nullptr);
cudaError_t launch_oneshot(const PeerTable& pt, const Ctl& ctl, uint64_t in_off, uint64_t out_off, uint64_t nvec,
                           int dtype, int max_ctas, float scale, cudaStream_t s);
// `scale` multiplies the reduced sum before it is written back (1.0 = none; the Reducer pre-scales in flatten
// like the reference, the comm-hook shim and in-place buckets post-scale here)
cudaError_t launch_twoshot(const PeerTable& pt, const Ctl& ctl, uint64_t in_off, uint64_t out_off, uint64_t nvec,
                           int dtype, int max_ctas, float scale, cudaStream_t s);
cudaError_t launch_twoshot_tma(const PeerTable& pt, const Ctl& ctl, uint64_t in_off, uint64_t out_off,
                               uint64_t nvec, int dtype, int max_ctas, float scale, cudaStream_t s);
cudaError_t launch_nvls(const PeerTable& pt, const Ctl& ctl, uint64_t mc_base, uint64_t in_off, uint64_t out_off,
                        uint64_t nvec, int dtype, int max_ctas, float scale, cudaStream_t s);
cudaError_t launch_fused(const TensorTable& tt, const PeerTable& pt, const Ctl& ctl, uint64_t bucket_off,
                         uint32_t nvec, int dtype, float scale, cudaStream_t s, const SgdTable* sgd = nullptr);
cudaError_t launch_broadcast(const PeerTable& pt, const Ctl& ctl, uint64_t off, uint64_t nvec, int root,
                             int max_ctas, cudaStream_t s);
cudaError_t launch_barrier(const PeerTable& pt, const Ctl& ctl, cudaStream_t s);
cudaError_t launch_local_copy(void* dst, const void* src, uint64_t nvec, int max_ctas, cudaStream_t s);
cudaError_t launch_gather_rows(const void* src, void* dst, void* dst2, uint32_t row_words, uint32_t split_words,
                               uint32_t batch_rows, uint64_t full_rows, void* cursor, cudaStream_t s);
cudaError_t prepare_kernels(int world);  // dynamic-smem opt-in + eager load of every kernel of this world size

}  // namespace b200ddp
