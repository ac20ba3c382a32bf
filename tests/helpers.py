"""Test-only helpers: engines that stand in for the CUDA engine so the Reducer's HOST logic can be
exercised on CPU (single process, or world_size>1 over gloo).  These live under tests/ on purpose: the
product package has no CPU implementation of the bucket engine."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

from b200ddp.bucketing import BucketSpec
from b200ddp.reducer import BucketEngine


class RecordingEngine(BucketEngine):
    """Single-rank engine that records the order of bucket launches and leaves gradients untouched."""

    def __init__(self, world_size: int = 1) -> None:
        self.rank, self.world_size = 0, world_size
        self.plans: List[List[BucketSpec]] = []
        self.launched: List[int] = []
        self.launch_params: List[List[int]] = []
        self.finalized = 0

    def plan(self, buckets):
        self.plans.append(list(buckets))

    def reduce_bucket(self, bucket, grads):
        assert all(g is not None for g in grads)
        self.launched.append(bucket.index)
        self.launch_params.append(list(bucket.param_indices))

    def finalize(self):
        self.finalized += 1

    def sync_order(self, order):
        return list(order)

    def broadcast_tensors(self, tensors, root=0):
        pass


class GlooEngine(BucketEngine):
    """world_size>1 on CPU: the bucket arithmetic is done by the ORACLE's definition (pre-scale by
    fl32(1/ws), rank-ordered fp32 sum) on top of gloo all_gather, so that a CPU run of our Reducer can be
    compared with stock DDP.  Checker only."""

    def __init__(self, sgd_epilogue: bool = False) -> None:
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        self.launched: List[int] = []
        if sgd_epilogue:  # stands in for the product engine's optimizer epilogue (DDP.fuse_optimizer)
            self.sgd_lr, self.params, self.epilogues = None, (), 0

    def plan(self, buckets):
        self.buckets = list(buckets)

    def reduce_bucket(self, bucket, grads):
        self.launched.append(bucket.index)
        scale = np.float32(1.0) / np.float32(self.world_size)
        flat = torch.cat([g.reshape(-1) for g in grads]) * float(scale)
        gathered = [torch.empty_like(flat) for _ in range(self.world_size)]
        dist.all_gather(gathered, flat)
        acc = gathered[0].clone()
        for t in gathered[1:]:
            acc += t
        off = 0
        for g in grads:
            g.copy_(acc[off:off + g.numel()].view_as(g))
            off += g.numel()
        if getattr(self, "sgd_lr", None) is not None:  # p <- p - lr g where the averaged gradient is written back
            with torch.no_grad():
                for i, g in zip(bucket.param_indices, grads):
                    self.params[i].add_(g, alpha=-self.sgd_lr)
            self.epilogues += 1

    def finalize(self):
        pass

    def sync_order(self, order):
        box = [order]
        dist.broadcast_object_list(box, src=0)
        return list(box[0])

    def broadcast_tensors(self, tensors: Sequence[torch.Tensor], root: int = 0):
        for t in tensors:
            dist.broadcast(t, src=root)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| — the parity metric of BASELINE.md §2."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    denom = float(b.abs().max())
    num = float((a - b).abs().max())
    return num / denom if denom > 0 else num
