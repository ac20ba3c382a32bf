"""Oracle restatement of c10d's ``compute_bucket_assignment_by_size`` — TEST INFRASTRUCTURE.

Follows the published algorithm of torch/csrc/distributed/c10d/reducer.cpp (declaration:
torch:include/torch/csrc/distributed/c10d/reducer.hpp:592-598; limits reducer.hpp:30-31; how DDP calls it:
torch:nn/parallel/distributed.py:1219-1247 and Reducer::rebuild_buckets reducer.hpp:136).  Written
independently of distributed-pytorch_b200/bucketing.py and pinned against torch's own function through
tests/golden/bucket_assignment.json.

Rules restated:
  R1  tensors are visited in the given order; tensor at position p has index p, or tensor_indices[p];
  R2  there is one accumulating bucket per (dtype, device) key;
  R3  each key walks the list of size limits independently, starting at the first limit;
  R4  after appending a tensor, if the accumulated byte size >= the key's current limit the bucket is
      emitted with that limit, the accumulator restarts empty, and the key advances to the next limit
      unless it already is at the last one;
  R5  at the end every non-empty accumulator is emitted with the limit it was last compared against
      (torch walks a std::unordered_map here, so with SEVERAL keys and ready-order input the order of these
      trailing buckets is implementation defined; this restatement emits them in first-seen key order —
      membership and limits are pinned by the golden vectors, see tests/test_oracle_golden.py);
  R6  if tensor_indices was NOT given, emitted buckets are ordered by their smallest tensor index;
      otherwise emission order is kept (gradient-ready order).
"""
from typing import Dict, Hashable, List, Optional, Sequence, Tuple


class _Accumulator:
    def __init__(self) -> None:
        self.members: List[int] = []
        self.bytes = 0
        self.limit_pos = 0
        self.last_limit = 0


def assign_buckets(
    nbytes: Sequence[int],
    keys: Sequence[Hashable],
    limits: Sequence[int],
    tensor_indices: Optional[Sequence[int]] = None,
) -> Tuple[List[List[int]], List[int]]:
    accs: Dict[Hashable, _Accumulator] = {}
    emitted: List[Tuple[List[int], int]] = []
    for position in range(len(nbytes)):
        index = position if not tensor_indices else tensor_indices[position]  # R1
        acc = accs.setdefault(keys[position], _Accumulator())  # R2, R3
        acc.members.append(index)
        acc.bytes += nbytes[position]
        acc.last_limit = limits[acc.limit_pos]
        if acc.bytes >= acc.last_limit:  # R4
            emitted.append((acc.members, acc.last_limit))
            acc.members = []
            acc.bytes = 0
            acc.limit_pos = min(acc.limit_pos + 1, len(limits) - 1)
    for acc in accs.values():  # R5 (dict preserves first-seen key order)
        if acc.members:
            emitted.append((acc.members, acc.last_limit))
    if not tensor_indices:  # R6
        emitted = sorted(emitted, key=lambda e: min(e[0]))
    return [m for m, _ in emitted], [l for _, l in emitted]


def ddp_initial_buckets(nbytes, keys, first_cap=1 << 20, cap=25 << 20):
    """Construction-time buckets as DDP hands them to its Reducer: assignment over parameters in
    definition order, list reversed (torch:nn/parallel/distributed.py:1240-1257)."""
    buckets, _ = assign_buckets(nbytes, keys, [first_cap, cap])
    return list(reversed(buckets))


def ddp_rebuilt_buckets(nbytes, keys, ready_order, first_cap=1 << 20, cap=25 << 20):
    """Buckets after the one-time rebuild: parameters in gradient-ready order (reducer.hpp:136)."""
    buckets, _ = assign_buckets([nbytes[i] for i in ready_order], [keys[i] for i in ready_order],
                                [first_cap, cap], tensor_indices=list(ready_order))
    return buckets
