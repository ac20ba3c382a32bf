"""``init_process_group`` / ``destroy_process_group`` with the reference's call signatures
(ref:multigpu.py:8,20,88; ref:multigpu_torchrun.py:8,13,94) — what ``ddp_setup()`` calls.

torch's process group is still created (it provides the TCP store rendezvous, ``DistributedSampler``'s
rank/world size and host-side barriers) but it carries no gradient traffic: the per-GPU arenas are mapped
peer-to-peer over NVLink when a ``DistributedDataParallel`` is constructed, and the allreduce kernels of
libb200ddp.so issue the transfers themselves.
"""
from __future__ import annotations

import os
from typing import Any, Optional

import torch
import torch.distributed as dist


def init_process_group(backend: Optional[str] = "nccl", *args: Any, **kwargs: Any) -> None:
    """Same arguments as ``torch.distributed.init_process_group``.  Additionally binds this process to
    its GPU (LOCAL_RANK under torchrun, else rank modulo the visible device count) so that NCCL's lazy
    communicator and our arena land on the right device."""
    dist.init_process_group(backend, *args, **kwargs)
    if torch.cuda.is_available():
        local = os.environ.get("LOCAL_RANK")
        dev = int(local) if local is not None else dist.get_rank() % torch.cuda.device_count()
        torch.cuda.set_device(dev)


def destroy_process_group(group: Any = None) -> None:
    """Tear down our communicators (fenced by a host barrier so no peer still reads our arena), then
    torch's process group."""
    from .ddp import _live_ddp, shutdown_all

    if dist.is_initialized():
        if _live_ddp and dist.get_world_size() > 1:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dist.barrier()
        shutdown_all()
        dist.destroy_process_group(group)
    else:
        shutdown_all()
