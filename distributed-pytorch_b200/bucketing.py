"""Gradient bucket assignment and arena layout (host-side integer bookkeeping of the path).

``compute_bucket_assignment_by_size`` follows c10d's function of the same name
(torch:include/torch/csrc/distributed/c10d/reducer.hpp:592-598; caps ``kDefaultFirstBucketBytes`` = 1 MiB
and ``kDefaultBucketBytesCap`` = 25 MiB at reducer.hpp:30-31; called from
torch:nn/parallel/distributed.py:1240 at construction and by ``Reducer::rebuild_buckets``,
reducer.hpp:136, after the first iteration).  The oracle (oracle/bucket_oracle.py) restates the same rules
independently and is pinned against torch's own function in tests/golden/.

``BucketLayout`` is ours: it decides where each bucket lives in the IPC-shared arena.  Every tensor
starts on a 16-byte boundary inside its bucket so the flatten kernel's bucket-side accesses are always
full 16-byte vectors (the reference packs tensors back to back, reducer.hpp:374-375; the padding never
leaves the arena, so results are unaffected).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

DEFAULT_FIRST_BUCKET_BYTES = 1024 * 1024  # reducer.hpp:30
DEFAULT_BUCKET_BYTES_CAP = 25 * 1024 * 1024  # reducer.hpp:31

VEC_BYTES = 16


def compute_bucket_assignment_by_size(
    sizes_bytes: Sequence[int],
    keys: Sequence[object],
    bucket_size_limits: Sequence[int],
    tensor_indices: Optional[Sequence[int]] = None,
) -> Tuple[List[List[int]], List[int]]:
    """Greedy size-capped bucketing, one open bucket per key (dtype, device).

    ``sizes_bytes[i]`` / ``keys[i]`` describe tensor i in the order given.  A tensor is appended to the
    open bucket of its key; when that bucket reaches the key's current limit it is closed and the key
    moves on to the next limit (the last limit repeats).  Left-over open buckets are appended at the end.
    Without ``tensor_indices`` the result is sorted by each bucket's smallest tensor index; with it the
    tensors are already in gradient-ready order, positions are mapped through it and nothing is sorted.
    Returns (bucket_indices, per_bucket_size_limits).
    """
    if len(sizes_bytes) != len(keys):
        raise ValueError("sizes_bytes and keys differ in length")
    if tensor_indices is not None and len(tensor_indices) != len(sizes_bytes):
        raise ValueError("tensor_indices must name every tensor")
    if not bucket_size_limits:
        raise ValueError("need at least one bucket size limit")
    limits = [int(x) for x in bucket_size_limits]
    open_buckets: dict = {}  # key -> [indices, size, limit_position]
    key_order: List[object] = []
    result: List[Tuple[List[int], int]] = []
    for pos, (nbytes, key) in enumerate(zip(sizes_bytes, keys)):
        idx = pos if tensor_indices is None else int(tensor_indices[pos])
        if key not in open_buckets:
            open_buckets[key] = [[], 0, 0]
            key_order.append(key)
        acc = open_buckets[key]
        acc[0].append(idx)
        acc[1] += int(nbytes)
        limit = limits[acc[2]]
        if acc[1] >= limit:
            result.append((acc[0], limit))
            acc[0], acc[1] = [], 0
            if acc[2] + 1 < len(limits):
                acc[2] += 1
    for key in key_order:
        acc = open_buckets[key]
        if acc[0]:
            result.append((acc[0], limits[acc[2]]))
    if tensor_indices is None:
        result.sort(key=lambda b: min(b[0]))
    return [b[0] for b in result], [b[1] for b in result]


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


@dataclass
class BucketSpec:
    """One gradient bucket: which parameters, where they sit inside it, where it sits in the arena."""

    index: int
    param_indices: List[int]
    dtype_code: int  # _lib.F32 / BF16 / F16
    elem_size: int
    offsets_el: List[int] = field(default_factory=list)  # per parameter, in elements, 16-byte aligned
    lengths_el: List[int] = field(default_factory=list)
    nbytes: int = 0  # padded bucket size, multiple of 16
    arena_off: int = 0  # where peers read this rank's contribution
    out_off: int = 0  # where the reduced bucket lands (== arena_off for in-place algorithms)
    algo: int = 0
    fused: bool = False

    @property
    def payload_bytes(self) -> int:
        return sum(self.lengths_el) * self.elem_size


def layout_bucket(index: int, param_indices: Sequence[int], numels: Sequence[int], dtype_code: int,
                  elem_size: int) -> BucketSpec:
    """Place the parameters of one bucket back to back, each padded to a 16-byte boundary."""
    epv = VEC_BYTES // elem_size
    spec = BucketSpec(index=index, param_indices=list(param_indices), dtype_code=dtype_code, elem_size=elem_size)
    off = 0
    for n in numels:
        spec.offsets_el.append(off)
        spec.lengths_el.append(int(n))
        off += _round_up(int(n), epv)
    spec.nbytes = max(off * elem_size, VEC_BYTES)
    return spec


class BucketLayout:
    """Assigns arena regions to buckets from ``start`` on.  Out-of-place (one-shot) and fused (double
    buffered) buckets get two consecutive regions."""

    REGION_ALIGN = 4096

    def __init__(self, start: int = 0) -> None:
        self.total_bytes = _round_up(int(start), self.REGION_ALIGN)

    def place(self, spec: BucketSpec, out_of_place: bool = False, region_bytes: Optional[int] = None) -> None:
        """``region_bytes``: reserve exactly this much (the fused path's double-buffered / LL region)."""
        spec.arena_off = self.total_bytes
        if region_bytes is not None:
            self.total_bytes += _round_up(max(region_bytes, spec.nbytes), self.REGION_ALIGN)
            spec.out_off = spec.arena_off
            return
        self.total_bytes += _round_up(spec.nbytes, self.REGION_ALIGN)
        if out_of_place:
            spec.out_off = self.total_bytes
            self.total_bytes += _round_up(spec.nbytes, self.REGION_ALIGN)
        else:
            spec.out_off = spec.arena_off
