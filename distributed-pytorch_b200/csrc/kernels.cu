// kernels.cu — sm_100a kernels of the B200-native DDP gradient-bucket path.
//
// What each kernel replaces in the reference's dependency stack (torch c10d, see include/b200ddp.h):
//   multi_tensor_copy_kernel  Reducer::mark_variable_ready_dense (flatten + 1/ws pre-scale, reducer.hpp:276)
//                             and copy_bucket_to_grad (unflatten, reducer.hpp:330): one launch per bucket
//                             instead of one D2D launch per parameter.
//   oneshot/twoshot/tma/nvls  ProcessGroupNCCL::allreduce -> ncclAllReduce(SUM) (ProcessGroupNCCL.hpp:856):
//                             the kernel itself issues the NVLink traffic on peer-mapped arenas.
//   fused_bucket_kernel       all three steps of a small bucket in a single launch (one cross-GPU barrier);
//   fused_ll_kernel           the same for buckets <= 32 KiB with NO barrier (data + tag in one 8-byte store).
//   gather_rows_kernel        the per-batch `source.to(gpu_id)` of ref:multigpu.py:48-50 for a shard resident in HBM.
//   (SGD instantiations)      fused_*/unflatten with an SgdTable also apply torch.optim.SGD's p -= lr g (opt-in).
//
// Design notes
//   * Everything is a stream of 16-byte vectors (LDG.128/STG.128, L1::no_allocate): the path is pure
//     bandwidth (0.25 FLOP/B), no tensor cores.
//   * Sums are accumulated in fp32 in RANK ORDER 0..ws-1 on exactly one owner per element (two-shot) or
//     identically on every rank (one-shot), so all ranks end with bit-identical buckets and the fp32
//     result is independent of algorithm and grid size (NVLS excepted: the switch picks the order).
//   * Cross-GPU synchronisation is a per-CTA flag barrier on peer-mapped signal pads: CTA b of rank r
//     stores a monotonically increasing epoch into slot [b][r] of every peer with st.release.sys and
//     spins with ld.acquire.sys on its own pad.  The epoch lives in device memory (per CTA), so a
//     launch carries no host-side generation counter and can be replayed from a CUDA graph.
//     CTA b only ever depends on CTA b of the peers, so no grid-wide co-residency is needed.
//   * Spins are bounded by %globaltimer; a timeout raises a flag in mapped host memory instead of
//     hanging the GPU (a dead peer becomes a Python exception) and poisons the rank: the kernel returns
//     without touching gradients, later collective kernels return at once, unflatten leaves .grad alone.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "internal.h"

namespace b200ddp {
namespace {

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------------------
// 16-byte global accesses
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_vec(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_vec(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds_vec(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}

// ---------------------------------------------------------------------------------------------------
// element-type arithmetic on one 16-byte vector; accumulation always in fp32, never contracted
// ---------------------------------------------------------------------------------------------------
template <int DT>
struct Vec;

template <>
struct Vec<B200DDP_F32> {
  static constexpr int N = 4;      // elements per 16-byte vector
  static constexpr int ESIZE = 4;  // bytes per element
  static __device__ __forceinline__ void unpack(const uint4& x, float (&f)[N]) {
    f[0] = __uint_as_float(x.x);
    f[1] = __uint_as_float(x.y);
    f[2] = __uint_as_float(x.z);
    f[3] = __uint_as_float(x.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[N]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};

template <>
struct Vec<B200DDP_BF16> {
  static constexpr int N = 8;
  static constexpr int ESIZE = 2;
  static __device__ __forceinline__ void unpack(const uint4& x, float (&f)[N]) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[N]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <>
struct Vec<B200DDP_F16> {
  static constexpr int N = 8;
  static constexpr int ESIZE = 2;
  static __device__ __forceinline__ void unpack(const uint4& x, float (&f)[N]) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = p.x;
      f[2 * i + 1] = p.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[N]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// sum of WS vectors in index order 0..WS-1 (= rank order), fp32 accumulate, one rounding on pack
template <int DT, int WS>
__device__ __forceinline__ uint4 reduce_vecs(const uint4 (&x)[WS]) {
  float acc[Vec<DT>::N];
  Vec<DT>::unpack(x[0], acc);
#pragma unroll
  for (int q = 1; q < WS; ++q) {
    float f[Vec<DT>::N];
    Vec<DT>::unpack(x[q], f);
#pragma unroll
    for (int i = 0; i < Vec<DT>::N; ++i) acc[i] = __fadd_rn(acc[i], f[i]);
  }
  return Vec<DT>::pack(acc);
}

template <int DT>
__device__ __forceinline__ uint4 scale_vec(const uint4& x, float scale) {
  // scale == 1 is a bit-exact pass-through: the same kernels move raw bytes of any dtype (parameter and
  // buffer broadcast), where a multiply would canonicalise NaN payloads of reinterpreted integers
  if (scale == 1.0f) return x;
  float f[Vec<DT>::N];
  Vec<DT>::unpack(x, f);
#pragma unroll
  for (int i = 0; i < Vec<DT>::N; ++i) f[i] = __fmul_rn(f[i], scale);
  return Vec<DT>::pack(f);
}

// ---------------------------------------------------------------------------------------------------
// cross-GPU per-CTA barrier on the signal pads
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t* flag_slot(uint64_t base, int block, int src) {
  return reinterpret_cast<uint32_t*>(base) + (block * kMaxWorld + src);
}
__device__ __forceinline__ uint32_t* epoch_slot(uint64_t base, int block) {
  return reinterpret_cast<uint32_t*>(base + kFlagBytes) + block;
}
__device__ __forceinline__ uint32_t read_epoch(const PeerTable& pt, const Ctl& ctl) {
  return *reinterpret_cast<volatile uint32_t*>(epoch_slot(pt.base[ctl.rank], blockIdx.x));
}
__device__ __forceinline__ void write_epoch(const PeerTable& pt, const Ctl& ctl, uint32_t e) {
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(epoch_slot(pt.base[ctl.rank], blockIdx.x)) = e;
}

// Sticky per-rank "poisoned" word (device memory, local signal region): set by the first kernel of this rank
// that times out on a peer.  Later collective kernels return at once and unflatten launches leave the
// gradients untouched, so a half-reduced bucket is never scattered into .grad; the host sees the mapped error
// word (b200ddp_poll_error) and raises.
__device__ __forceinline__ uint32_t* poison_slot(uint64_t base) {
  return reinterpret_cast<uint32_t*>(base + kFlagBytes + kEpochBytes);
}
__device__ __forceinline__ bool poisoned(const PeerTable& pt, const Ctl& ctl) {
  return *reinterpret_cast<volatile uint32_t*>(poison_slot(pt.base[ctl.rank])) != 0u;
}
__device__ __forceinline__ void raise_timeout(const PeerTable& pt, const Ctl& ctl, int peer) {
  *reinterpret_cast<volatile uint32_t*>(poison_slot(pt.base[ctl.rank])) = 1u;
  *reinterpret_cast<volatile uint32_t*>(ctl.err) = 1u + static_cast<uint32_t>(peer);
}

// Every thread of the CTA calls this.  On return (true) all CTAs with this blockIdx on all ranks have arrived
// with `target`, and the writes (local and peer) every thread of those CTAs made before arriving are
// visible to this CTA.  Ordering: the CTA's writes happen-before the bar.sync, the per-peer thread's
// st.release.sys is cumulative over them (PTX memory model, causality order through the CTA barrier), the
// waiter's ld.acquire.sys + bar.sync extends that to every thread here — the same pattern as the
// block barriers of NCCL's and TensorRT-LLM's one/two-shot kernels.
// Returns false (to every thread of the CTA) when a peer did not arrive within the timeout: the error word in
// mapped host memory is set and the caller must NOT touch gradients or peer data any more (a poisoned
// bucket is left as it was instead of being overwritten with a partial sum).
__device__ __forceinline__ bool cta_barrier(const PeerTable& pt, const Ctl& ctl, int ws, uint32_t target) {
  __syncthreads();
  int ok = 1;
  if (threadIdx.x < ws) {
    const int q = threadIdx.x;
    st_release_sys(flag_slot(pt.base[q], blockIdx.x, ctl.rank), target);
    const uint32_t* mine = flag_slot(pt.base[ctl.rank], blockIdx.x, q);
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - target) < 0) {
      if ((++spins & 255u) == 0) {  // the timer is only consulted every 256 polls
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > ctl.timeout_ns) {
          raise_timeout(pt, ctl, q);
          ok = 0;
          break;
        }
      }
    }
  }
  return __syncthreads_and(ok) != 0;
}

__device__ __forceinline__ uint64_t arena_of(const PeerTable& pt, int q) { return pt.base[q] + kSignalBytes; }

// ---------------------------------------------------------------------------------------------------
// multi-tensor flatten(+scale) / unflatten
// ---------------------------------------------------------------------------------------------------
struct SmemTable {
  uint64_t ptr[kMaxTableTensors];
  uint64_t len[kMaxTableTensors];
  uint32_t off[kMaxTableTensors + 1];
};

__device__ __forceinline__ void load_table(SmemTable& s, const TensorTable& tt) {
  const int n = tt.n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s.ptr[i] = tt.ptr[i];
    s.len[i] = tt.len_el[i];
  }
  for (int i = threadIdx.x; i <= n; i += blockDim.x) s.off[i] = tt.off_vec[i];
  __syncthreads();
}

// largest t in [lo, hi] with off[t] <= v   (requires off[lo] <= v)
__device__ __forceinline__ int locate(const uint32_t* off, int lo, int hi, uint32_t v) {
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= v)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// read bucket-vector v of tensor t from the tensor's own storage (zero padded past its end)
template <int DT>
__device__ __forceinline__ uint4 gather_vec(const SmemTable& s, int t, uint32_t v) {
  constexpr int N = Vec<DT>::N;
  const uint64_t e0 = static_cast<uint64_t>(v - s.off[t]) * N;
  const uint64_t len = s.len[t];
  const uint64_t addr = s.ptr[t] + e0 * Vec<DT>::ESIZE;
  if (e0 + N <= len && (addr & 15) == 0) return ld_vec(reinterpret_cast<const void*>(addr));
  uint4 r = make_uint4(0, 0, 0, 0);
  if (e0 >= len) return r;
  const int valid = static_cast<int>(min(static_cast<uint64_t>(N), len - e0));
  uint32_t w[4] = {0, 0, 0, 0};
  if (Vec<DT>::ESIZE == 4) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(addr);
    for (int i = 0; i < valid; ++i) w[i] = p[i];
  } else {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(addr);
    for (int i = 0; i < valid; ++i) w[i >> 1] |= static_cast<uint32_t>(p[i]) << (16 * (i & 1));
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// write bucket-vector v of tensor t to the tensor's own storage (only its valid elements)
template <int DT>
__device__ __forceinline__ void scatter_vec(const SmemTable& s, int t, uint32_t v, const uint4& x) {
  constexpr int N = Vec<DT>::N;
  const uint64_t e0 = static_cast<uint64_t>(v - s.off[t]) * N;
  const uint64_t len = s.len[t];
  if (e0 >= len) return;
  const uint64_t addr = s.ptr[t] + e0 * Vec<DT>::ESIZE;
  if (e0 + N <= len && (addr & 15) == 0) {
    st_vec(reinterpret_cast<void*>(addr), x);
    return;
  }
  const int valid = static_cast<int>(min(static_cast<uint64_t>(N), len - e0));
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
  if (Vec<DT>::ESIZE == 4) {
    uint32_t* p = reinterpret_cast<uint32_t*>(addr);
    for (int i = 0; i < valid; ++i) p[i] = w[i];
  } else {
    uint16_t* p = reinterpret_cast<uint16_t*>(addr);
    for (int i = 0; i < valid; ++i) p[i] = static_cast<uint16_t>(w[i >> 1] >> (16 * (i & 1)));
  }
}

// ---- optional optimizer epilogue (opt-in, SgdTable in internal.h) -----------------------------------------------
struct NoSgd {};
template <bool SGD>
struct SgdArg {
  using type = NoSgd;
};
template <>
struct SgdArg<true> {
  using type = SgdTable;
};

// p <- fma(neg_lr, g, p) on the elements of tensor storage `pbase` (len elements) that bucket vector element e0.. covers
template <int DT>
__device__ __forceinline__ void sgd_vec(uint64_t pbase, uint64_t len, uint64_t e0, const uint4& g, float neg_lr) {
  constexpr int N = Vec<DT>::N;
  if (pbase == 0 || e0 >= len) return;
  const uint64_t addr = pbase + e0 * Vec<DT>::ESIZE;
  float gf[N];
  Vec<DT>::unpack(g, gf);
  if (e0 + N <= len && (addr & 15) == 0) {
    float pf[N];
    Vec<DT>::unpack(ld_vec(reinterpret_cast<const void*>(addr)), pf);
#pragma unroll
    for (int i = 0; i < N; ++i) pf[i] = __fmaf_rn(neg_lr, gf[i], pf[i]);
    st_vec(reinterpret_cast<void*>(addr), Vec<DT>::pack(pf));
    return;
  }
  const int valid = static_cast<int>(min(static_cast<uint64_t>(N), len - e0));
  if (Vec<DT>::ESIZE == 4) {
    float* p = reinterpret_cast<float*>(addr);
    for (int i = 0; i < valid; ++i) p[i] = __fmaf_rn(neg_lr, gf[i], p[i]);
  } else {
    uint16_t* p = reinterpret_cast<uint16_t*>(addr);
    for (int i = 0; i < valid; ++i) {
      float one[N] = {};
      Vec<DT>::unpack(make_uint4(p[i], 0, 0, 0), one);
      one[0] = __fmaf_rn(neg_lr, gf[i], one[0]);
      p[i] = static_cast<uint16_t>(Vec<DT>::pack(one).x);
    }
  }
}

constexpr int kCopyThreads = 512;

// One launch moves bucket vectors [vec_begin, vec_end) (all inside the tensors of `tt`).
// HBM-bound: reads S + writes S bytes for S bytes of gradients.
// A tile is kCopyThreads * UNROLL vectors; every thread keeps UNROLL 16-byte loads in flight before its first
// store.  Tiles that lie inside one aligned tensor take a straight vector copy; tiles that span tensor
// boundaries (ResNet-50: 107 of 161 gradients are < 16 KiB), tails or unaligned storage look the tensor of every
// vector up in the shared-memory table (binary search over the <= few tensors of the tile) but keep the same
// loads-then-stores structure, so such a tile costs one memory round trip like any other (round 1 walked it
// one dependent vector at a time and those tiles finished last: 0.44 of the HBM roofline on ResNet-50 buckets).
template <int DT, bool UNFLATTEN, int UNROLL, bool SGD = false>
__global__ void __launch_bounds__(kCopyThreads, (UNROLL <= 4 && !SGD) ? 4 : 2)
multi_tensor_copy_kernel(const __grid_constant__ TensorTable tt, unsigned char* __restrict__ bucket,
                         uint32_t vec_begin, uint32_t vec_end, float scale, int even_split,
                         const uint32_t* __restrict__ poison, const __grid_constant__ typename SgdArg<SGD>::type sgd) {
  static_assert(!SGD || UNFLATTEN, "the optimizer epilogue belongs to the unflatten direction");
  constexpr uint32_t kTile = kCopyThreads * UNROLL;
  __shared__ SmemTable s;
  // unflatten after a peer timeout: the bucket may hold a partial sum, leave the gradients alone
  if (UNFLATTEN && poison != nullptr && *reinterpret_cast<const volatile uint32_t*>(poison) != 0u) return;
  // The tile-level lookups below are CTA-uniform and read the table straight from the kernel-parameter constant
  // bank (uniform index: one broadcast LDC); the shared-memory copy, which per-THREAD lookups need, is only made by
  // a CTA that meets a boundary tile.  A CTA whose tiles all lie inside aligned tensors has no prologue at all.
  bool table_in_smem = false;
  const int n = tt.n;
  // Default grid: one CTA per tile (launch_copy_variant); with fewer CTAs than tiles the tiles are dealt
  // round-robin (the CTAs sweep memory as a sliding window).  even_split: every CTA owns one contiguous, equal
  // share instead — no tail, but measurably worse DRAM locality (see launch_flatten).
  const uint32_t granules = (vec_end - vec_begin + kCopyThreads - 1) / kCopyThreads;
  const uint32_t per = granules / gridDim.x, extra = granules % gridDim.x;
  const uint32_t g0 = blockIdx.x * per + min(blockIdx.x, extra);
  const uint32_t g1 = g0 + per + (blockIdx.x < extra ? 1u : 0u);
  const uint32_t r0 = even_split ? vec_begin + g0 * kCopyThreads : vec_begin + blockIdx.x * kTile;
  const uint32_t r1 = even_split ? min(vec_end, vec_begin + g1 * kCopyThreads) : vec_end;
  const uint32_t stride = even_split ? kTile : gridDim.x * kTile;
  uint4* bvec = reinterpret_cast<uint4*>(bucket);
  for (uint32_t v0 = r0; v0 < r1; v0 += stride) {
    const uint32_t v1 = min(v0 + kTile, r1);
    const int t_lo = locate(tt.off_vec, 0, n - 1, v0);
    const int t_hi = locate(tt.off_vec, t_lo, n - 1, v1 - 1);
    const uint64_t first = static_cast<uint64_t>(v0 - tt.off_vec[t_lo]);
    bool fast = (t_lo == t_hi) && ((tt.ptr[t_lo] & 15) == 0) &&
                (static_cast<uint64_t>(v1 - tt.off_vec[t_lo]) * Vec<DT>::N <= tt.len_el[t_lo]);
    if constexpr (SGD) fast = fast && ((sgd.ptr[t_lo] & 15) == 0);  // an unaligned parameter takes the general path
    uint4 x[UNROLL];
    if (!fast && !table_in_smem) {  // CTA-uniform branch
      load_table(s, tt);
      table_in_smem = true;
    }
    if (fast) {
      // whole tile inside one aligned tensor: straight vector copy
      uint4* tvec = reinterpret_cast<uint4*>(tt.ptr[t_lo]) + first;
      const uint32_t cnt = v1 - v0;
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t i = threadIdx.x + j * kCopyThreads;
        if (i < cnt) x[j] = UNFLATTEN ? ld_vec(bvec + v0 + i) : ld_vec(tvec + i);
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t i = threadIdx.x + j * kCopyThreads;
        if (i < cnt) {
          if (UNFLATTEN)
            st_vec(tvec + i, x[j]);
          else
            st_vec(bvec + v0 + i, scale_vec<DT>(x[j], scale));
        }
      }
      if constexpr (SGD) {
        // p <- fma(-lr, g, p) on the same vectors (the parameter is laid out like its gradient)
        if (sgd.ptr[t_lo] != 0) {
          uint4* pvec = reinterpret_cast<uint4*>(sgd.ptr[t_lo]) + first;
          uint4 p[UNROLL];
#pragma unroll
          for (int j = 0; j < UNROLL; ++j) {
            const uint32_t i = threadIdx.x + j * kCopyThreads;
            if (i < cnt) p[j] = ld_vec(pvec + i);
          }
#pragma unroll
          for (int j = 0; j < UNROLL; ++j) {
            const uint32_t i = threadIdx.x + j * kCopyThreads;
            if (i < cnt) {
              float pf[Vec<DT>::N], gf[Vec<DT>::N];
              Vec<DT>::unpack(p[j], pf);
              Vec<DT>::unpack(x[j], gf);
#pragma unroll
              for (int k = 0; k < Vec<DT>::N; ++k) pf[k] = __fmaf_rn(sgd.neg_lr, gf[k], pf[k]);
              st_vec(pvec + i, Vec<DT>::pack(pf));
            }
          }
        }
      }
    } else if (UNFLATTEN) {
      // bucket-side loads do not depend on the tensor lookup: issue them all, search while they fly
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t v = v0 + threadIdx.x + j * kCopyThreads;
        if (v < v1) x[j] = ld_vec(bvec + v);
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t v = v0 + threadIdx.x + j * kCopyThreads;
        if (v < v1) {
          const int t = locate(s.off, t_lo, t_hi, v);
          scatter_vec<DT>(s, t, v, x[j]);
          if constexpr (SGD)
            sgd_vec<DT>(sgd.ptr[t], s.len[t], static_cast<uint64_t>(v - s.off[t]) * Vec<DT>::N, x[j], sgd.neg_lr);
        }
      }
    } else {
      int t[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t v = v0 + threadIdx.x + j * kCopyThreads;
        t[j] = v < v1 ? locate(s.off, t_lo, t_hi, v) : t_lo;
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t v = v0 + threadIdx.x + j * kCopyThreads;
        if (v < v1) x[j] = gather_vec<DT>(s, t[j], v);
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const uint32_t v = v0 + threadIdx.x + j * kCopyThreads;
        if (v < v1) st_vec(bvec + v, scale_vec<DT>(x[j], scale));
      }
    }
  }
}

// plain vector copy, 4 loads in flight per thread (world-1 out-of-place "allreduce", and the peer-read
// bandwidth probe: src may be a peer's mapping, then the loads cross NVLink)
__global__ void __launch_bounds__(512) local_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src,
                                                         uint64_t nvec) {
  constexpr int U = 4;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x * U;
  for (uint64_t i0 = blockIdx.x * uint64_t(blockDim.x) * U + threadIdx.x; i0 < nvec; i0 += step) {
    uint4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + uint64_t(u) * blockDim.x < nvec) x[u] = ld_vec(src + i0 + uint64_t(u) * blockDim.x);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + uint64_t(u) * blockDim.x < nvec) st_vec(dst + i0 + uint64_t(u) * blockDim.x, x[u]);
  }
}

// ---------------------------------------------------------------------------------------------------
// input pipeline: the next batch of a device-resident shard, gathered by a device-side cursor.
// dst[b] = src[(cursor + b) mod full_rows] for b < batch_rows (rows of row_words 4-byte words), then
// cursor = (cursor + batch_rows) mod full_rows.  One launch, no host state: a captured training step replays it.
// Single CTA: every thread reads the cursor before the barrier, thread 0 advances it after.
// ---------------------------------------------------------------------------------------------------
// split_words > 0: the first split_words words of every row go to dst as a dense [batch, split_words] block, the rest
// to dst2 as a dense [batch, row_words - split_words] block (source and targets land contiguous, no strided views).
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                          uint32_t* __restrict__ dst2, uint32_t row_words,
                                                          uint32_t split_words, uint32_t batch_rows, uint64_t full_rows,
                                                          unsigned long long* cursor) {
  const uint64_t cur = *reinterpret_cast<volatile unsigned long long*>(cursor) % full_rows;
  __syncthreads();
  const uint32_t total = batch_rows * row_words;
  const uint32_t rest = row_words - split_words;
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
    const uint32_t b = i / row_words, k = i - b * row_words;
    const uint32_t v = src[((cur + b) % full_rows) * row_words + k];
    if (split_words == 0)
      dst[i] = v;
    else if (k < split_words)
      dst[b * split_words + k] = v;
    else
      dst2[b * rest + (k - split_words)] = v;
  }
  if (threadIdx.x == 0) *cursor = (cur + batch_rows) % full_rows;
}

// ---------------------------------------------------------------------------------------------------
// one-shot allreduce: every rank reads the whole bucket from every peer (small buckets, latency bound)
// ---------------------------------------------------------------------------------------------------
constexpr int kArThreads = 512;

template <int DT, int WS>
__global__ void __launch_bounds__(kArThreads)
oneshot_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ Ctl ctl, uint64_t in_off,
               uint64_t out_off, uint64_t nvec, float scale) {
  if (poisoned(pt, ctl)) return;
  const uint32_t e = read_epoch(pt, ctl);
  if (!cta_barrier(pt, ctl, WS, e + 1)) return;
  const uint4* in[WS];
#pragma unroll
  for (int q = 0; q < WS; ++q) in[q] = reinterpret_cast<const uint4*>(arena_of(pt, q) + in_off);
  uint4* out = reinterpret_cast<uint4*>(arena_of(pt, ctl.rank) + out_off);
  for (uint64_t i = blockIdx.x * uint64_t(kArThreads) + threadIdx.x; i < nvec; i += uint64_t(gridDim.x) * kArThreads) {
    uint4 x[WS];
#pragma unroll
    for (int q = 0; q < WS; ++q) x[q] = ld_vec(in[q] + i);
    st_vec(out + i, scale_vec<DT>(reduce_vecs<DT, WS>(x), scale));
  }
  cta_barrier(pt, ctl, WS, e + 2);
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// two-shot allreduce, register path: rank r owns slice r; it loads that slice from all ws arenas with
// LDG.128, reduces, and pushes the result into all ws arenas with STG.128 (reduce-scatter and
// all-gather fused per vector; in-place safe because the owner is the only reader of its slice).
// NVLink bytes per GPU and direction: 2 (ws-1)/ws S.
// ---------------------------------------------------------------------------------------------------
constexpr int kTwoShotUnroll = 2;

template <int DT, int WS>
__global__ void __launch_bounds__(kArThreads)
twoshot_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ Ctl ctl, uint64_t in_off,
               uint64_t out_off, uint64_t nvec, float scale) {
  if (poisoned(pt, ctl)) return;
  const uint32_t e = read_epoch(pt, ctl);
  if (!cta_barrier(pt, ctl, WS, e + 1)) return;
  const uint64_t slice = (nvec + WS - 1) / WS;
  const uint64_t begin = min(nvec, slice * ctl.rank);
  const uint64_t end = min(nvec, begin + slice);
  const uint4* in[WS];
  uint4* out[WS];
#pragma unroll
  for (int q = 0; q < WS; ++q) {
    in[q] = reinterpret_cast<const uint4*>(arena_of(pt, q) + in_off);
    out[q] = reinterpret_cast<uint4*>(arena_of(pt, q) + out_off);
  }
  constexpr uint64_t kStep = uint64_t(kArThreads) * kTwoShotUnroll;
  for (uint64_t i0 = begin + blockIdx.x * kStep + threadIdx.x; i0 < end; i0 += uint64_t(gridDim.x) * kStep) {
    uint4 x[kTwoShotUnroll][WS];
#pragma unroll
    for (int u = 0; u < kTwoShotUnroll; ++u) {
      const uint64_t i = i0 + uint64_t(u) * kArThreads;
      if (i < end) {
#pragma unroll
        for (int q = 0; q < WS; ++q) x[u][q] = ld_vec(in[q] + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kTwoShotUnroll; ++u) {
      const uint64_t i = i0 + uint64_t(u) * kArThreads;
      if (i < end) {
        const uint4 r = scale_vec<DT>(reduce_vecs<DT, WS>(x[u]), scale);
#pragma unroll
        for (int q = 0; q < WS; ++q) st_vec(out[q] + i, r);
      }
    }
  }
  cta_barrier(pt, ctl, WS, e + 2);
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// two-shot allreduce, TMA path: one elected producer thread streams 4 KiB pieces of the owned slice
// from all ws arenas into a shared-memory ring with cp.async.bulk (UBLKCP) completing on mbarriers;
// 8 consumer warps reduce from shared memory and push results to all arenas.  Bytes in flight per CTA
// = 128 KiB independent of register count, so a few CTAs saturate NVLink and the backward keeps the SMs.
// ---------------------------------------------------------------------------------------------------
constexpr int kTmaChunkVec = 256;                       // vectors per arena per stage (4 KiB)
constexpr int kTmaChunkBytes = kTmaChunkVec * 16;       //
constexpr int kTmaConsumerWarps = 8;                    // 256 consumer threads = one vector each per stage
constexpr int kTmaThreads = 32 + 32 * kTmaConsumerWarps;
__host__ __device__ constexpr int tma_stages(int ws) { return 32 / ws; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int DT, int WS>
__global__ void __launch_bounds__(kTmaThreads)
twoshot_tma_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ Ctl ctl, uint64_t in_off,
                   uint64_t out_off, uint64_t nvec, float scale) {
  if (poisoned(pt, ctl)) return;
  constexpr int STAGES = tma_stages(WS);
  extern __shared__ __align__(128) unsigned char ring[];  // STAGES * WS * 4 KiB
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
  const uint32_t full_s = static_cast<uint32_t>(__cvta_generic_to_shared(full_bar));
  const uint32_t empty_s = static_cast<uint32_t>(__cvta_generic_to_shared(empty_bar));
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_s + 8 * i, 1);
      mbar_init(empty_s + 8 * i, kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t e = read_epoch(pt, ctl);
  if (!cta_barrier(pt, ctl, WS, e + 1)) return;  // also publishes the mbarrier inits (contains __syncthreads)

  const uint64_t slice = (nvec + WS - 1) / WS;
  const uint64_t begin = min(nvec, slice * ctl.rank);
  const uint64_t end = min(nvec, begin + slice);
  const uint64_t ntiles = (end - begin + kTmaChunkVec - 1) / kTmaChunkVec;

  if (warp == 0) {
    if (lane == 0) {
      // the peers' flatten writes were observed through the generic proxy (ld.acquire.sys in cta_barrier); the
      // bulk copies below read them through the async proxy: order the two proxies explicitly
      asm volatile("fence.proxy.async.global;" ::: "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(empty_s + 8 * stage, phase ^ 1);
        const uint64_t v0 = begin + t * kTmaChunkVec;
        const uint32_t bytes = static_cast<uint32_t>(min(static_cast<uint64_t>(kTmaChunkVec), end - v0)) * 16;
        mbar_expect_tx(full_s + 8 * stage, bytes * WS);
#pragma unroll
        for (int q = 0; q < WS; ++q)
          bulk_g2s(ring_s + (stage * WS + q) * kTmaChunkBytes,
                   reinterpret_cast<const void*>(arena_of(pt, q) + in_off + v0 * 16), bytes, full_s + 8 * stage);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    const int ctid = tid - 32;
    uint4* out[WS];
#pragma unroll
    for (int q = 0; q < WS; ++q) out[q] = reinterpret_cast<uint4*>(arena_of(pt, q) + out_off);
    int stage = 0;
    uint32_t phase = 0;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(full_s + 8 * stage, phase);
      const uint64_t v0 = begin + t * kTmaChunkVec;
      const uint32_t cnt = static_cast<uint32_t>(min(static_cast<uint64_t>(kTmaChunkVec), end - v0));
      if (ctid < cnt) {
        uint4 x[WS];
#pragma unroll
        for (int q = 0; q < WS; ++q) x[q] = lds_vec(ring_s + (stage * WS + q) * kTmaChunkBytes + ctid * 16);
        const uint4 r = scale_vec<DT>(reduce_vecs<DT, WS>(x), scale);
#pragma unroll
        for (int q = 0; q < WS; ++q) st_vec(out[q] + v0 + ctid, r);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + 8 * stage);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
  cta_barrier(pt, ctl, WS, e + 2);
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// NVLS allreduce: the owner of a slice pulls the switch-reduced sum with multimem.ld_reduce and
// broadcasts it with multimem.st through the multicast mapping.  NVLink bytes per GPU and direction:
// (1 + 1/ws) S instead of 2 (ws-1)/ws S.
// ---------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ uint4 multimem_ld_reduce(const void* p);
template <>
__device__ __forceinline__ uint4 multimem_ld_reduce<B200DDP_F32>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 multimem_ld_reduce<B200DDP_BF16>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 multimem_ld_reduce<B200DDP_F16>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(void* p, const uint4& v) {
  // the payload is opaque 16 bytes; .f32 only names the register class
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

template <int DT, int kNvlsUnroll>
__global__ void __launch_bounds__(kArThreads)
nvls_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ Ctl ctl, uint64_t mc_base, uint64_t in_off,
            uint64_t out_off, uint64_t nvec, float scale) {
  if (poisoned(pt, ctl)) return;
  const int ws = ctl.world;
  const uint32_t e = read_epoch(pt, ctl);
  if (!cta_barrier(pt, ctl, ws, e + 1)) return;
  const uint64_t slice = (nvec + ws - 1) / ws;
  const uint64_t begin = min(nvec, slice * ctl.rank);
  const uint64_t end = min(nvec, begin + slice);
  const uint4* in = reinterpret_cast<const uint4*>(mc_base + kSignalBytes + in_off);
  uint4* out = reinterpret_cast<uint4*>(mc_base + kSignalBytes + out_off);
  constexpr uint64_t kStep = uint64_t(kArThreads) * kNvlsUnroll;
  for (uint64_t i0 = begin + blockIdx.x * kStep + threadIdx.x; i0 < end; i0 += uint64_t(gridDim.x) * kStep) {
    uint4 x[kNvlsUnroll];
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const uint64_t i = i0 + uint64_t(u) * kArThreads;
      if (i < end) x[u] = multimem_ld_reduce<DT>(in + i);
    }
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const uint64_t i = i0 + uint64_t(u) * kArThreads;
      if (i < end) multimem_st(out + i, scale_vec<DT>(x[u], scale));
    }
  }
  cta_barrier(pt, ctl, ws, e + 2);
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// fused small bucket: flatten+scale -> one-shot allreduce -> unflatten in ONE launch.
// CTA b owns bucket vectors [b*kFusedVecPerCta, ...): it gathers them from the gradients into its own
// arena, meets CTA b of every peer, sums the same range of all arenas and scatters the result straight
// back into the gradients from registers.
// The arena staging is DOUBLE BUFFERED (slot = bit 1 of the CTA's barrier generation), which removes the
// trailing barrier: a slot is only rewritten two generations later, and passing the start barrier of the
// generation in between proves every peer has finished reading it.  One NVLink round trip per bucket.
// ---------------------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;
constexpr uint32_t kFusedVecPerCta = 1024;

template <int DT, int WS, bool SGD = false>
__global__ void __launch_bounds__(kFusedThreads)
fused_bucket_kernel(const __grid_constant__ TensorTable tt, const __grid_constant__ PeerTable pt,
                    const __grid_constant__ Ctl ctl, uint64_t bucket_off, uint32_t nvec, float scale,
                    const __grid_constant__ typename SgdArg<SGD>::type sgd) {
  __shared__ SmemTable s;
  load_table(s, tt);
  const int n = tt.n;
  const uint32_t v_begin = blockIdx.x * kFusedVecPerCta;
  const uint32_t v_end = min(nvec, v_begin + kFusedVecPerCta);
  if (WS == 1) {
    uint4* mine = reinterpret_cast<uint4*>(arena_of(pt, ctl.rank) + bucket_off);
    for (uint32_t v = v_begin + threadIdx.x; v < v_end; v += kFusedThreads) {
      const int t = locate(s.off, 0, n - 1, v);
      const uint4 x = scale_vec<DT>(gather_vec<DT>(s, t, v), scale);
      st_vec(mine + v, x);
      scatter_vec<DT>(s, t, v, x);
      if constexpr (SGD)
        sgd_vec<DT>(sgd.ptr[t], s.len[t], static_cast<uint64_t>(v - s.off[t]) * Vec<DT>::N, x, sgd.neg_lr);
    }
    return;
  }
  if (poisoned(pt, ctl)) return;
  const uint32_t e = read_epoch(pt, ctl);
  const uint64_t slot_off = bucket_off + ((e >> 1) & 1u) * (static_cast<uint64_t>(nvec) * 16);
  uint4* mine = reinterpret_cast<uint4*>(arena_of(pt, ctl.rank) + slot_off);
  for (uint32_t v = v_begin + threadIdx.x; v < v_end; v += kFusedThreads) {
    const int t = locate(s.off, 0, n - 1, v);
    st_vec(mine + v, scale_vec<DT>(gather_vec<DT>(s, t, v), scale));
  }
  if (!cta_barrier(pt, ctl, WS, e + 1)) return;  // gradients keep their local values
  const uint4* in[WS];
#pragma unroll
  for (int q = 0; q < WS; ++q) in[q] = reinterpret_cast<const uint4*>(arena_of(pt, q) + slot_off);
  for (uint32_t v = v_begin + threadIdx.x; v < v_end; v += kFusedThreads) {
    uint4 x[WS];
#pragma unroll
    for (int q = 0; q < WS; ++q) x[q] = ld_vec(in[q] + v);
    const int t = locate(s.off, 0, n - 1, v);
    const uint4 r = reduce_vecs<DT, WS>(x);
    scatter_vec<DT>(s, t, v, r);
    if constexpr (SGD)
      sgd_vec<DT>(sgd.ptr[t], s.len[t], static_cast<uint64_t>(v - s.off[t]) * Vec<DT>::N, r, sgd.neg_lr);
  }
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// fused tiny bucket, LL protocol (<= kLLMaxBytes): NO barrier at all.  Every 4-byte word of the scaled
// bucket travels with a 4-byte tag in one 8-byte store (single-copy atomic over NVLink): each rank pushes
// its words into slot [set][rank] of every peer's LL region and then polls its own slots until all ws
// tags match, sums in rank order and scatters to the gradients.  Latency = one NVLink write + the poll,
// instead of barrier + read round trip + barrier.  The tag is the CTA's barrier generation (device
// memory, monotonically increasing, so stale slots can never match); two alternating slot sets make the
// rewrite of a slot safe (reaching generation g+2 needs every peer's data of g+1, which a peer only sends
// after it has consumed generation g).
// ---------------------------------------------------------------------------------------------------
constexpr int kLLThreads = 256;
constexpr uint32_t kLLWordsPerCta = 1024;

// One 64-bit scalar access per (data, tag) pair: single-copy atomic by the PTX memory model (an aligned
// scalar access is never torn), so a reader that sees the tag also sees the data it travelled with.
__device__ __forceinline__ void st_ll(uint2* p, uint32_t data, uint32_t tag) {
  const uint64_t packed = static_cast<uint64_t>(data) | (static_cast<uint64_t>(tag) << 32);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(packed) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
  uint64_t packed;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(packed) : "l"(p) : "memory");
  return make_uint2(static_cast<uint32_t>(packed), static_cast<uint32_t>(packed >> 32));
}

// bucket word w (4 bytes) of tensor t, read from the tensor's storage and pre-scaled
template <int DT>
__device__ __forceinline__ uint32_t gather_word(const SmemTable& s, int t, uint32_t w, float scale) {
  const uint32_t v = w >> 2, j = w & 3;
  const uint64_t len = s.len[t];
  if (Vec<DT>::ESIZE == 4) {
    const uint64_t e = static_cast<uint64_t>(v - s.off[t]) * 4 + j;
    if (e >= len) return 0u;
    const float x = reinterpret_cast<const float*>(s.ptr[t])[e];
    return __float_as_uint(scale == 1.0f ? x : __fmul_rn(x, scale));
  } else {
    const uint64_t e = static_cast<uint64_t>(v - s.off[t]) * 8 + 2 * j;
    const uint16_t* p = reinterpret_cast<const uint16_t*>(s.ptr[t]);
    uint32_t bits = 0;
    if (e < len) bits |= p[e];
    if (e + 1 < len) bits |= static_cast<uint32_t>(p[e + 1]) << 16;
    if (scale == 1.0f) return bits;
    // run the pair through the same unpack / multiply / round-once path as the vector kernels
    uint4 x = make_uint4(bits, 0, 0, 0);
    return scale_vec<DT>(x, scale).x;
  }
}

template <int DT>
__device__ __forceinline__ void scatter_word(const SmemTable& s, int t, uint32_t w, uint32_t bits) {
  const uint32_t v = w >> 2, j = w & 3;
  const uint64_t len = s.len[t];
  if (Vec<DT>::ESIZE == 4) {
    const uint64_t e = static_cast<uint64_t>(v - s.off[t]) * 4 + j;
    if (e < len) reinterpret_cast<uint32_t*>(s.ptr[t])[e] = bits;
  } else {
    const uint64_t e = static_cast<uint64_t>(v - s.off[t]) * 8 + 2 * j;
    uint16_t* p = reinterpret_cast<uint16_t*>(s.ptr[t]);
    if (e < len) p[e] = static_cast<uint16_t>(bits);
    if (e + 1 < len) p[e + 1] = static_cast<uint16_t>(bits >> 16);
  }
}

// p <- fma(neg_lr, g, p) for the one (fp32) or two (bf16/f16) elements bucket word w of tensor t covers
template <int DT>
__device__ __forceinline__ void sgd_word(uint64_t pbase, const SmemTable& s, int t, uint32_t w, uint32_t bits,
                                         float neg_lr) {
  if (pbase == 0) return;
  const uint32_t v = w >> 2, j = w & 3;
  const uint64_t len = s.len[t];
  float gf[Vec<DT>::N] = {};
  Vec<DT>::unpack(make_uint4(bits, 0, 0, 0), gf);
  if (Vec<DT>::ESIZE == 4) {
    const uint64_t e = static_cast<uint64_t>(v - s.off[t]) * 4 + j;
    if (e < len) {
      float* p = reinterpret_cast<float*>(pbase) + e;
      *p = __fmaf_rn(neg_lr, gf[0], *p);
    }
  } else {
    const uint64_t e = static_cast<uint64_t>(v - s.off[t]) * 8 + 2 * j;
    uint16_t* p = reinterpret_cast<uint16_t*>(pbase);
    for (int k = 0; k < 2; ++k) {
      if (e + k < len) {
        float one[Vec<DT>::N] = {};
        Vec<DT>::unpack(make_uint4(p[e + k], 0, 0, 0), one);
        one[0] = __fmaf_rn(neg_lr, gf[k], one[0]);
        p[e + k] = static_cast<uint16_t>(Vec<DT>::pack(one).x);
      }
    }
  }
}

template <int DT, int WS, bool SGD = false>
__global__ void __launch_bounds__(kLLThreads)
fused_ll_kernel(const __grid_constant__ TensorTable tt, const __grid_constant__ PeerTable pt,
                const __grid_constant__ Ctl ctl, uint64_t ll_off, uint32_t nwords, float scale,
                const __grid_constant__ typename SgdArg<SGD>::type sgd) {
  __shared__ SmemTable s;
  if (poisoned(pt, ctl)) return;
  const uint32_t e = read_epoch(pt, ctl);  // before the barrier inside load_table: thread 0 rewrites it at the end
  load_table(s, tt);
  const int n = tt.n;
  const uint32_t tag = e + 1;
  const uint64_t set_base = static_cast<uint64_t>((e >> 1) & 1u) * WS * nwords;  // in uint2 units
  const uint32_t w_begin = blockIdx.x * kLLWordsPerCta;
  const uint32_t w_end = min(nwords, w_begin + kLLWordsPerCta);
  // push: my words -> slot [set][my rank] on every rank (my own included)
  for (uint32_t w = w_begin + threadIdx.x; w < w_end; w += kLLThreads) {
    const int t = locate(s.off, 0, n - 1, w >> 2);
    const uint32_t d = gather_word<DT>(s, t, w, scale);
    const uint64_t slot = set_base + static_cast<uint64_t>(ctl.rank) * nwords + w;
#pragma unroll
    for (int q = 0; q < WS; ++q) st_ll(reinterpret_cast<uint2*>(arena_of(pt, q) + ll_off) + slot, d, tag);
  }
  // pull: poll my own slots of every source rank until all ws tags match this generation; the data words that
  // travelled with the tags stay in registers (a thread owns at most kLLWordsPerCta / kLLThreads words)
  constexpr int kPerThread = kLLWordsPerCta / kLLThreads;
  const uint2* mine = reinterpret_cast<const uint2*>(arena_of(pt, ctl.rank) + ll_off) + set_base;
  uint32_t val[kPerThread][WS];
  int timed_out = 0;
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    const uint32_t w = w_begin + threadIdx.x + k * kLLThreads;
    if (w < w_end && !timed_out) {
      uint64_t t0 = 0;
      uint32_t spins = 0;
      // one poll round = WS independent 8-byte loads in flight at once (round 1 polled the sources one after
      // the other: ws dependent L2 round trips, ~3 us of the 5 us this kernel took at ws = 8)
      for (;;) {
        uint2 r[WS];
#pragma unroll
        for (int q = 0; q < WS; ++q) r[q] = ld_ll(mine + static_cast<uint64_t>(q) * nwords + w);
        int missing = -1;
#pragma unroll
        for (int q = 0; q < WS; ++q) {
          val[k][q] = r[q].x;
          if (r[q].y != tag) missing = q;
        }
        if (missing < 0) break;
        if ((++spins & 255u) == 0) {
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          if (now - t0 > ctl.timeout_ns) {
            raise_timeout(pt, ctl, missing);
            timed_out = 1;
            break;
          }
        }
      }
    }
  }
  // a bucket that did not arrive completely is not scattered at all: the gradients keep their local values
  if (__syncthreads_or(timed_out)) return;
  // sum in rank order, scatter
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    const uint32_t w = w_begin + threadIdx.x + k * kLLThreads;
    if (w < w_end) {
      uint4 x[WS];
#pragma unroll
      for (int q = 0; q < WS; ++q) x[q] = make_uint4(val[k][q], 0, 0, 0);
      const int t = locate(s.off, 0, n - 1, w >> 2);
      const uint32_t bits = reduce_vecs<DT, WS>(x).x;
      scatter_word<DT>(s, t, w, bits);
      if constexpr (SGD) sgd_word<DT>(sgd.ptr[t], s, t, w, bits, sgd.neg_lr);
    }
  }
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// broadcast by peer stores, and a bare barrier
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kArThreads)
broadcast_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ Ctl ctl, uint64_t off, uint64_t nvec,
                 int root) {
  if (poisoned(pt, ctl)) return;
  const int ws = ctl.world;
  const uint32_t e = read_epoch(pt, ctl);
  if (!cta_barrier(pt, ctl, ws, e + 1)) return;
  if (ctl.rank == root) {
    const uint4* src = reinterpret_cast<const uint4*>(arena_of(pt, root) + off);
    for (uint64_t i = blockIdx.x * uint64_t(kArThreads) + threadIdx.x; i < nvec;
         i += uint64_t(gridDim.x) * kArThreads) {
      const uint4 x = ld_vec(src + i);
      for (int q = 0; q < ws; ++q)
        if (q != root) st_vec(reinterpret_cast<uint4*>(arena_of(pt, q) + off) + i, x);
    }
  }
  cta_barrier(pt, ctl, ws, e + 2);
  write_epoch(pt, ctl, e + 2);
}

__global__ void __launch_bounds__(32) barrier_kernel(const __grid_constant__ PeerTable pt,
                                                     const __grid_constant__ Ctl ctl) {
  if (poisoned(pt, ctl)) return;
  const uint32_t e = read_epoch(pt, ctl);
  cta_barrier(pt, ctl, ctl.world, e + 1);
  write_epoch(pt, ctl, e + 2);
}

// ---------------------------------------------------------------------------------------------------
// dispatch helpers
// ---------------------------------------------------------------------------------------------------
inline int clamp_grid(uint64_t want, int max_ctas, int dflt) {
  int cap = max_ctas > 0 ? max_ctas : dflt;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (want < 1) want = 1;
  return static_cast<int>(want < static_cast<uint64_t>(cap) ? want : static_cast<uint64_t>(cap));
}

#define B200_DISPATCH_DT(dt, ...)                       \
  switch (dt) {                                         \
    case B200DDP_F32: {                                 \
      constexpr int DT = B200DDP_F32;                   \
      __VA_ARGS__;                                      \
    } break;                                            \
    case B200DDP_BF16: {                                \
      constexpr int DT = B200DDP_BF16;                  \
      __VA_ARGS__;                                      \
    } break;                                            \
    case B200DDP_F16: {                                 \
      constexpr int DT = B200DDP_F16;                   \
      __VA_ARGS__;                                      \
    } break;                                            \
    default:                                            \
      return cudaErrorInvalidValue;                     \
  }

#define B200_WS_CASE(n, ...) \
  case n: {                  \
    constexpr int WS = n;    \
    __VA_ARGS__;             \
  } break;

#define B200_DISPATCH_WS(ws, ...)   \
  switch (ws) {                     \
    B200_WS_CASE(2, __VA_ARGS__)    \
    B200_WS_CASE(3, __VA_ARGS__)    \
    B200_WS_CASE(4, __VA_ARGS__)    \
    B200_WS_CASE(5, __VA_ARGS__)    \
    B200_WS_CASE(6, __VA_ARGS__)    \
    B200_WS_CASE(7, __VA_ARGS__)    \
    B200_WS_CASE(8, __VA_ARGS__)    \
    default:                        \
      return cudaErrorInvalidValue; \
  }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------------
namespace {
// One CTA per 32 KiB tile (up to 64 waves of 592 CTAs; beyond that CTAs loop): measured against a single co-resident
// wave of looping CTAs, tools/ab_copy.py on B200 (profiles/r02_copy_kernel_ab.txt): 256 MiB 6.55 vs 5.86 TB/s (1.00 vs
// 0.89 of the measured copy peak; torch's own copy kernel: 0.98), 64 MiB 0.89 vs 0.82-0.84, <= 30 MiB equal (the
// bucket is one or two waves either way).  The hardware CTA scheduler refills an SM as soon as a CTA's stores are
// issued, so loads of new tiles overlap the drain of old ones — what a looping CTA serialises.  Possible because a
// CTA whose tile lies inside one tensor has no prologue (table lookups from the constant bank, see the kernel).
constexpr uint32_t kCopyWavesDefault = 64;
// Vectors per thread and tile of the copy kernel (B200DDP_COPY_UNROLL = 2 | 4 | 8 forces one).  Measured on B200
// (tools/ab_copy.py, profiles/): 4 x 16 B in flight per thread at 4 CTAs x 512 threads per SM.
int copy_unroll() {
  const char* e = getenv("B200DDP_COPY_UNROLL");  // read per launch (like B200DDP_COPY_SPLIT): tools/ab_copy.py flips it
  const int v = e ? atoi(e) : 4;
  return (v == 2 || v == 8) ? v : 4;
}

template <int DT, int UNROLL>
void launch_copy_variant(const TensorTable& tt, unsigned char* b, uint32_t vec_begin, uint32_t vec_end, float scale,
                         bool unflatten, int even, const uint32_t* poison, cudaStream_t s) {
  const NoSgd none{};
  // one co-resident wave: 4 CTAs of 512 threads per SM up to UNROLL 4 (32 registers), 2 per SM at UNROLL 8;
  // B200DDP_COPY_WAVES = k launches up to k waves (k >= 64: one CTA per tile, the hardware scheduler refills SMs)
  const uint32_t tiles = (vec_end - vec_begin + kCopyThreads * UNROLL - 1) / (kCopyThreads * UNROLL);
  const char* we = getenv("B200DDP_COPY_WAVES");
  const uint32_t waves = we ? static_cast<uint32_t>(atoi(we) > 0 ? atoi(we) : 1) : kCopyWavesDefault;
  const uint32_t wave = uint32_t(kNumSMs) * (UNROLL <= 4 ? 4 : 2) * waves;
  const int grid = static_cast<int>(tiles < wave ? (tiles ? tiles : 1) : wave);
  if (unflatten)
    multi_tensor_copy_kernel<DT, true, UNROLL>
        <<<grid, kCopyThreads, 0, s>>>(tt, b, vec_begin, vec_end, scale, even, poison, none);
  else
    multi_tensor_copy_kernel<DT, false, UNROLL>
        <<<grid, kCopyThreads, 0, s>>>(tt, b, vec_begin, vec_end, scale, even, nullptr, none);
}
}  // namespace

cudaError_t launch_flatten(const TensorTable& tt, void* bucket, uint32_t vec_begin, uint32_t vec_end, int dtype,
                           float scale, bool unflatten, const uint32_t* poison, cudaStream_t s, const SgdTable* sgd) {
  if (vec_end <= vec_begin || tt.n <= 0) return cudaSuccess;
  unsigned char* b = static_cast<unsigned char*>(bucket);
  if (sgd != nullptr) {
    // unflatten + optimizer epilogue: one CTA per 32 KiB tile, 2 CTAs per SM (the parameter vectors double the
    // registers in flight)
    if (!unflatten) return cudaErrorInvalidValue;
    const uint32_t tiles = (vec_end - vec_begin + kCopyThreads * 4 - 1) / (kCopyThreads * 4);
    const int grid = static_cast<int>(tiles < 65535u ? tiles : 65535u);
    B200_DISPATCH_DT(dtype, {
      multi_tensor_copy_kernel<DT, true, 4, true>
          <<<grid, kCopyThreads, 0, s>>>(tt, b, vec_begin, vec_end, scale, 0, poison, *sgd);
    });
    return cudaGetLastError();
  }
  // Work split, A/B-measured on one box (tools/ab_copy.py): dealing tiles round-robin makes the co-resident
  // CTAs sweep memory as a sliding window and beat an even per-CTA split by 2-7 % at 16-256 MiB (25 MiB:
  // 4.88 vs 4.58 TB/s), so it is the default; B200DDP_COPY_SPLIT=even selects the other.
  const char* mode = getenv("B200DDP_COPY_SPLIT");
  const int even = (mode && mode[0] == 'e') ? 1 : 0;
  const int unroll = copy_unroll();
  B200_DISPATCH_DT(dtype, {
    if (unroll == 2)
      launch_copy_variant<DT, 2>(tt, b, vec_begin, vec_end, scale, unflatten, even, poison, s);
    else if (unroll == 8)
      launch_copy_variant<DT, 8>(tt, b, vec_begin, vec_end, scale, unflatten, even, poison, s);
    else
      launch_copy_variant<DT, 4>(tt, b, vec_begin, vec_end, scale, unflatten, even, poison, s);
  });
  return cudaGetLastError();
}

cudaError_t launch_local_copy(void* dst, const void* src, uint64_t nvec, int max_ctas, cudaStream_t s) {
  if (nvec == 0) return cudaSuccess;
  const int grid = clamp_grid((nvec + 2047) / 2048, max_ctas, kNumSMs * 4);
  local_copy_kernel<<<grid, 512, 0, s>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), nvec);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const void* src, void* dst, void* dst2, uint32_t row_words, uint32_t split_words,
                               uint32_t batch_rows, uint64_t full_rows, void* cursor, cudaStream_t s) {
  gather_rows_kernel<<<1, 256, 0, s>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst),
                                       static_cast<uint32_t*>(dst2), row_words, split_words, batch_rows, full_rows,
                                       static_cast<unsigned long long*>(cursor));
  return cudaGetLastError();
}

cudaError_t launch_oneshot(const PeerTable& pt, const Ctl& ctl, uint64_t in_off, uint64_t out_off, uint64_t nvec,
                           int dtype, int max_ctas, float scale, cudaStream_t s) {
  const int grid = clamp_grid((nvec + kArThreads - 1) / kArThreads, max_ctas, 36);
  B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                     oneshot_kernel<DT, WS><<<grid, kArThreads, 0, s>>>(pt, ctl, in_off, out_off, nvec, scale);
                   }));
  return cudaGetLastError();
}

cudaError_t launch_twoshot(const PeerTable& pt, const Ctl& ctl, uint64_t in_off, uint64_t out_off, uint64_t nvec,
                           int dtype, int max_ctas, float scale, cudaStream_t s) {
  const uint64_t slice = (nvec + ctl.world - 1) / ctl.world;
  const uint64_t step = uint64_t(kArThreads) * kTwoShotUnroll;
  const int grid = clamp_grid((slice + step - 1) / step, max_ctas, 72);
  B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                     twoshot_kernel<DT, WS><<<grid, kArThreads, 0, s>>>(pt, ctl, in_off, out_off, nvec, scale);
                   }));
  return cudaGetLastError();
}

cudaError_t launch_twoshot_tma(const PeerTable& pt, const Ctl& ctl, uint64_t in_off, uint64_t out_off,
                               uint64_t nvec, int dtype, int max_ctas, float scale, cudaStream_t s) {
  const uint64_t slice = (nvec + ctl.world - 1) / ctl.world;
  const int grid = clamp_grid((slice + kTmaChunkVec - 1) / kTmaChunkVec, max_ctas, 36);
  B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                     constexpr size_t smem = size_t(tma_stages(WS)) * WS * kTmaChunkBytes;
                     twoshot_tma_kernel<DT, WS><<<grid, kTmaThreads, smem, s>>>(pt, ctl, in_off, out_off, nvec, scale);
                   }));
  return cudaGetLastError();
}

cudaError_t launch_nvls(const PeerTable& pt, const Ctl& ctl, uint64_t mc_base, uint64_t in_off, uint64_t out_off,
                        uint64_t nvec, int dtype, int max_ctas, float scale, cudaStream_t s) {
  const uint64_t slice = (nvec + ctl.world - 1) / ctl.world;
  static const int unroll = [] {
    const char* e = getenv("B200DDP_NVLS_UNROLL");
    const int u = e ? atoi(e) : 4;
    return (u == 2 || u == 8) ? u : 4;
  }();
  const uint64_t step = uint64_t(kArThreads) * unroll;
  // measured (profiles/r01_allreduce_sweep.md): at ws = 8, 8 CTAs x 512 threads x 4 vectors in flight is the
  // sweet spot at every size (more CTAs only add contention at the switch) and leaves 140 SMs to the
  // backward; a rank's slice is S/ws, so smaller worlds get proportionally more CTAs (ws = 2: 8 CTAs
  // measured 210 GB/s against 393 GB/s with 36)
  const int dflt = ctl.world >= 8 ? 8 : (ctl.world >= 4 ? 16 : 36);
  const int grid = clamp_grid((slice + step - 1) / step, max_ctas, dflt);
  B200_DISPATCH_DT(dtype, {
    if (unroll == 2)
      nvls_kernel<DT, 2><<<grid, kArThreads, 0, s>>>(pt, ctl, mc_base, in_off, out_off, nvec, scale);
    else if (unroll == 8)
      nvls_kernel<DT, 8><<<grid, kArThreads, 0, s>>>(pt, ctl, mc_base, in_off, out_off, nvec, scale);
    else
      nvls_kernel<DT, 4><<<grid, kArThreads, 0, s>>>(pt, ctl, mc_base, in_off, out_off, nvec, scale);
  });
  return cudaGetLastError();
}

cudaError_t launch_fused(const TensorTable& tt, const PeerTable& pt, const Ctl& ctl, uint64_t bucket_off,
                         uint32_t nvec, int dtype, float scale, cudaStream_t s, const SgdTable* sgd) {
  if (nvec == 0 || tt.n <= 0) return cudaSuccess;
  const int grid = static_cast<int>((nvec + kFusedVecPerCta - 1) / kFusedVecPerCta);
  if (grid > kMaxBlocks) return cudaErrorInvalidValue;
  const NoSgd none{};
  if (sgd != nullptr) {  // the same three kernels with the optimizer epilogue
    if (ctl.world == 1) {
      B200_DISPATCH_DT(dtype, {
        fused_bucket_kernel<DT, 1, true><<<grid, kFusedThreads, 0, s>>>(tt, pt, ctl, bucket_off, nvec, scale, *sgd);
      });
    } else if (uint64_t(nvec) * 16 <= kLLMaxBytes) {
      const uint32_t nwords = nvec * 4;
      const int ll_grid = static_cast<int>((nwords + kLLWordsPerCta - 1) / kLLWordsPerCta);
      B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                         fused_ll_kernel<DT, WS, true>
                             <<<ll_grid, kLLThreads, 0, s>>>(tt, pt, ctl, bucket_off, nwords, scale, *sgd);
                       }));
    } else {
      B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                         fused_bucket_kernel<DT, WS, true>
                             <<<grid, kFusedThreads, 0, s>>>(tt, pt, ctl, bucket_off, nvec, scale, *sgd);
                       }));
    }
    return cudaGetLastError();
  }
  if (ctl.world == 1) {
    B200_DISPATCH_DT(dtype, {
      fused_bucket_kernel<DT, 1><<<grid, kFusedThreads, 0, s>>>(tt, pt, ctl, bucket_off, nvec, scale, none);
    });
  } else if (uint64_t(nvec) * 16 <= kLLMaxBytes) {
    const uint32_t nwords = nvec * 4;
    const int ll_grid = static_cast<int>((nwords + kLLWordsPerCta - 1) / kLLWordsPerCta);
    B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                       fused_ll_kernel<DT, WS>
                           <<<ll_grid, kLLThreads, 0, s>>>(tt, pt, ctl, bucket_off, nwords, scale, none);
                     }));
  } else {
    B200_DISPATCH_DT(dtype, B200_DISPATCH_WS(ctl.world, {
                       fused_bucket_kernel<DT, WS>
                           <<<grid, kFusedThreads, 0, s>>>(tt, pt, ctl, bucket_off, nvec, scale, none);
                     }));
  }
  return cudaGetLastError();
}

cudaError_t launch_broadcast(const PeerTable& pt, const Ctl& ctl, uint64_t off, uint64_t nvec, int root,
                             int max_ctas, cudaStream_t s) {
  const int grid = clamp_grid((nvec + kArThreads * 4 - 1) / (kArThreads * 4), max_ctas, 36);
  broadcast_kernel<<<grid, kArThreads, 0, s>>>(pt, ctl, off, nvec, root);
  return cudaGetLastError();
}

cudaError_t launch_barrier(const PeerTable& pt, const Ctl& ctl, cudaStream_t s) {
  barrier_kernel<<<1, 32, 0, s>>>(pt, ctl);
  return cudaGetLastError();
}

namespace {
// CUDA loads kernels lazily, on first launch, and a load cannot complete while another kernel is still
// running on the device.  A kernel that spins for a peer must therefore never be the reason a peer-side
// launch is stuck behind a pending load (it deadlocks a single-GPU loop-back world outright, and is a
// needless serialisation point otherwise): every kernel this world size can launch is loaded up front.
template <typename K>
cudaError_t preload(K kernel) {
  cudaFuncAttributes attr;
  return cudaFuncGetAttributes(&attr, kernel);
}

template <int DT, int WS>
cudaError_t prepare_ws() {
  cudaError_t e;
  constexpr int smem = tma_stages(WS) * WS * kTmaChunkBytes;
  if ((e = cudaFuncSetAttribute(twoshot_tma_kernel<DT, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) !=
      cudaSuccess)
    return e;
  if ((e = preload(twoshot_tma_kernel<DT, WS>)) != cudaSuccess) return e;
  if ((e = preload(oneshot_kernel<DT, WS>)) != cudaSuccess) return e;
  if ((e = preload(twoshot_kernel<DT, WS>)) != cudaSuccess) return e;
  if ((e = preload(fused_bucket_kernel<DT, WS>)) != cudaSuccess) return e;
  if ((e = preload(fused_bucket_kernel<DT, WS, true>)) != cudaSuccess) return e;
  if ((e = preload(fused_ll_kernel<DT, WS, true>)) != cudaSuccess) return e;
  return preload(fused_ll_kernel<DT, WS>);
}

template <int DT>
cudaError_t prepare_dt(int world) {
  cudaError_t e;
  if ((e = preload(multi_tensor_copy_kernel<DT, false, 2>)) != cudaSuccess) return e;
  if ((e = preload(multi_tensor_copy_kernel<DT, true, 2>)) != cudaSuccess) return e;
  if ((e = preload(multi_tensor_copy_kernel<DT, false, 4>)) != cudaSuccess) return e;
  if ((e = preload(multi_tensor_copy_kernel<DT, true, 4>)) != cudaSuccess) return e;
  if ((e = preload(multi_tensor_copy_kernel<DT, false, 8>)) != cudaSuccess) return e;
  if ((e = preload(multi_tensor_copy_kernel<DT, true, 8>)) != cudaSuccess) return e;
  if ((e = preload(fused_bucket_kernel<DT, 1>)) != cudaSuccess) return e;
  if ((e = preload(fused_bucket_kernel<DT, 1, true>)) != cudaSuccess) return e;
  if ((e = preload(multi_tensor_copy_kernel<DT, true, 4, true>)) != cudaSuccess) return e;
  if ((e = preload(nvls_kernel<DT, 2>)) != cudaSuccess) return e;
  if ((e = preload(nvls_kernel<DT, 4>)) != cudaSuccess) return e;
  if ((e = preload(nvls_kernel<DT, 8>)) != cudaSuccess) return e;
  switch (world) {
    case 1: return cudaSuccess;
    case 2: return prepare_ws<DT, 2>();
    case 3: return prepare_ws<DT, 3>();
    case 4: return prepare_ws<DT, 4>();
    case 5: return prepare_ws<DT, 5>();
    case 6: return prepare_ws<DT, 6>();
    case 7: return prepare_ws<DT, 7>();
    case 8: return prepare_ws<DT, 8>();
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

cudaError_t prepare_kernels(int world) {
  cudaError_t e;
  if ((e = preload(local_copy_kernel)) != cudaSuccess) return e;
  if ((e = preload(gather_rows_kernel)) != cudaSuccess) return e;
  if ((e = preload(broadcast_kernel)) != cudaSuccess) return e;
  if ((e = preload(barrier_kernel)) != cudaSuccess) return e;
  if ((e = prepare_dt<B200DDP_F32>(world)) != cudaSuccess) return e;
  if ((e = prepare_dt<B200DDP_BF16>(world)) != cudaSuccess) return e;
  return prepare_dt<B200DDP_F16>(world);
}

}  // namespace b200ddp
