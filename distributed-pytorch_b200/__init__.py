"""b200ddp — a B200-native implementation of the one data-parallel hot path of
subramen/distributed-pytorch: the per-step gradient-bucket allreduce behind ``DDP(model)``.

Drop-in surface (the two import lines of the reference scripts, ref:multigpu.py:7-8)::

    from b200ddp import DistributedDataParallel as DDP
    from b200ddp import init_process_group, destroy_process_group

Everything below that surface is a custom Reducer (reducer.py) feeding hand-written sm_100a CUDA kernels
through a C ABI (include/b200ddp.h, csrc/).  There is no CPU, NCCL or eager-PyTorch fallback for the
allreduce: a missing library or a non-CUDA module raises.
"""
from .bucketing import compute_bucket_assignment_by_size  # noqa: F401
from .ddp import DistributedDataParallel  # noqa: F401
from .process_group import destroy_process_group, init_process_group  # noqa: F401

DDP = DistributedDataParallel
__all__ = [
    "DDP",
    "DistributedDataParallel",
    "init_process_group",
    "destroy_process_group",
    "compute_bucket_assignment_by_size",
]
