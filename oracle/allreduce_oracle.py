"""Oracle restatement of the bucket arithmetic — TEST INFRASTRUCTURE (numpy, CPU).

What the reference stack computes per bucket, per step:
  1. flatten + pre-scale  bucket[off_i : off_i+len_i] = grad_i * fl32(1/ws)
                          (Reducer::mark_variable_ready_dense, reducer.hpp:276; the multiply is fused into
                          the copy when no comm hook is installed, default_comm_hooks.hpp:36-41)
  2. allreduce(SUM)       bucket = sum over ranks (ProcessGroupNCCL::allreduce, ProcessGroupNCCL.hpp:856)
  3. unflatten            grad_i = bucket[off_i : off_i+len_i]   (copy_bucket_to_grad, reducer.hpp:330)

NCCL/gloo do not fix the order of the floating-point sum; our kernels do: rank order 0..ws-1, fp32
accumulation, one rounding to the storage type.  This oracle computes exactly that, so CUDA-vs-oracle
comparisons of fp32 buckets are BIT-EXACT, while oracle-vs-stock-DDP comparisons are within the
documented tolerance (max|a-b| / max|b| <= 1e-6 for fp32).
"""
from typing import List, Sequence

import numpy as np


# ---- bfloat16 / float16 storage helpers ------------------------------------------------------------
def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 (what cvt.rn.bf16x2.f32 does); NaN stays NaN."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        rounded = np.where(nan, np.uint16(0x7FFF), rounded)
    return rounded


def to_f32(storage: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "f32":
        return storage.astype(np.float32, copy=False)
    if dtype == "bf16":
        return bf16_bits_to_f32(storage)
    if dtype == "f16":
        return storage.view(np.float16).astype(np.float32)
    raise ValueError(dtype)


def from_f32(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "f32":
        return x.astype(np.float32, copy=False)
    if dtype == "bf16":
        return f32_to_bf16_bits(x)
    if dtype == "f16":
        return x.astype(np.float16).view(np.uint16)
    raise ValueError(dtype)


def storage_dtype(dtype: str):
    return np.float32 if dtype == "f32" else np.uint16


def elem_size(dtype: str) -> int:
    return 4 if dtype == "f32" else 2


# ---- the three steps --------------------------------------------------------------------------------
def padded_offsets(lengths: Sequence[int], dtype: str) -> List[int]:
    """Element offsets with every tensor starting on a 16-byte boundary (our arena layout)."""
    epv = 16 // elem_size(dtype)
    offs, off = [], 0
    for n in lengths:
        offs.append(off)
        off += (n + epv - 1) // epv * epv
    return offs


def flatten_scale(grads: Sequence[np.ndarray], offsets: Sequence[int], total_el: int, scale: float,
                  dtype: str) -> np.ndarray:
    """Step 1.  ``grads`` are 1-D storage arrays (float32, or uint16 bit patterns for bf16/f16)."""
    bucket = np.zeros(total_el, dtype=storage_dtype(dtype))
    s = np.float32(scale)
    for g, off in zip(grads, offsets):
        if dtype == "f32" and s == np.float32(1.0):
            bucket[off:off + g.size] = g
        else:
            bucket[off:off + g.size] = from_f32(to_f32(g, dtype) * s, dtype)
    return bucket


def allreduce_sum(buckets: Sequence[np.ndarray], dtype: str) -> np.ndarray:
    """Step 2, our definition: fp32 accumulation in rank order, one rounding at the end."""
    acc = to_f32(buckets[0], dtype).astype(np.float32, copy=True)
    for b in buckets[1:]:
        acc = (acc + to_f32(b, dtype)).astype(np.float32)
    return from_f32(acc, dtype)


def unflatten(bucket: np.ndarray, offsets: Sequence[int], lengths: Sequence[int]) -> List[np.ndarray]:
    """Step 3."""
    return [bucket[o:o + n].copy() for o, n in zip(offsets, lengths)]


def ddp_average(per_rank_grads: Sequence[Sequence[np.ndarray]], dtype: str = "f32") -> List[np.ndarray]:
    """All three steps for one bucket: per_rank_grads[r][i] is rank r's gradient of tensor i."""
    ws = len(per_rank_grads)
    lengths = [g.size for g in per_rank_grads[0]]
    offs = padded_offsets(lengths, dtype)
    epv = 16 // elem_size(dtype)
    total = offs[-1] + (lengths[-1] + epv - 1) // epv * epv if lengths else 0
    scale = float(np.float32(1.0) / np.float32(ws))
    flat = [flatten_scale([np.ravel(g) for g in grads], offs, total, scale, dtype) for grads in per_rank_grads]
    return unflatten(allreduce_sum(flat, dtype), offs, lengths)


# ---- round 2: the two optional arithmetic extras of the C ABI ------------------------------------------------
def allreduce_sum_scaled(buckets: Sequence[np.ndarray], dtype: str, scale: float) -> np.ndarray:
    """``b200ddp_allreduce(..., scale)``: the rank-ordered fp32 sum multiplied by fl32(scale) BEFORE the one
    rounding to the storage type — what the comm-hook shim uses, since torch hands a hook an UNdivided bucket
    (torch:nn/parallel/distributed.py:2021-2023; the default path divides first, default_comm_hooks.hpp:36-41).
    ``scale == 1`` is a bit-exact pass-through of the sum."""
    acc = to_f32(buckets[0], dtype).astype(np.float32, copy=True)
    for b in buckets[1:]:
        acc = (acc + to_f32(b, dtype)).astype(np.float32)
    s = np.float32(scale)
    if s == np.float32(1.0):
        return from_f32(acc, dtype)
    if dtype == "f32":
        return (acc * s).astype(np.float32)
    # 16-bit storage: the kernels round the sum to the storage type, then scale in fp32 and round again
    # (scale_vec on the packed vector): two roundings, restated literally
    summed = to_f32(from_f32(acc, dtype), dtype)
    return from_f32((summed * s).astype(np.float32), dtype)


def sgd_epilogue(param: np.ndarray, grad: np.ndarray, lr: float, dtype: str = "f32") -> np.ndarray:
    """The opt-in optimizer epilogue (``b200ddp_unflatten_sgd`` / ``b200ddp_fused_bucket_sgd``): plain SGD as
    ``torch.optim.SGD(lr)`` without momentum / weight decay applies it, ``param.add_(grad, alpha=-lr)``
    (ref:multigpu.py:43,68), computed as ONE fused multiply-add in fp32 and rounded once to the storage type:
    p <- fma(-lr, g, p).  The fma is restated in float64 (the product of two fp32 values is exact there)."""
    p64 = to_f32(param, dtype).astype(np.float64)
    g64 = to_f32(grad, dtype).astype(np.float64)
    exact = np.float64(np.float32(-lr)) * g64 + p64  # product exact, one float64 rounding on the sum
    return from_f32(exact.astype(np.float32), dtype)
