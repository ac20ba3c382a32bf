"""Comm-hook shim: our allreduce kernels behind STOCK ``torch.nn.parallel.DistributedDataParallel``.

torch's DDP has its own plugin point for the bucket -> allreduce step of the path,
``ddp.register_comm_hook(state, hook)`` with ``hook(state, bucket: dist.GradBucket) -> Future[Tensor]``
(torch:nn/parallel/distributed.py:1987-2067; C++ side ``CommHookInterface::runHook``, comm.hpp:104-118).  The
reference scripts bind ``DDP`` at ref:multigpu.py:7,36; a maintainer who wants to keep torch's Reducer and swap
ONLY the collective adds two lines after that wrap::

    from b200ddp import hooks
    self.model = DDP(model, device_ids=[gpu_id])                       # ref:multigpu.py:36, unchanged
    self.model.register_comm_hook(hooks.AllreduceState(), hooks.allreduce_hook)

This is the A/B harness of SURVEY §8(b): it isolates the kernel from our Reducer (torch still issues its
per-parameter copies into ``bucket.buffer()``).  The shipped path is the custom Reducer (ddp.py), which also
removes those copies.

Conventions of the hook contract honoured here: registered once, before the first backward
(distributed.py:2025-2027); the bucket is NOT pre-divided when a hook is installed (:2021-2023), so the kernel
multiplies the reduced sum by 1/world (``scale`` of ``b200ddp_allreduce``); the returned tensor is the bucket's own
flat buffer, same shape (:2029-2031); the second parameter is literally named ``bucket`` and annotated
``dist.GradBucket`` (``_check_comm_hook``, :2268-2289).
"""
# NOTE: no `from __future__ import annotations` here: torch's _check_comm_hook compares the hook's annotation
# OBJECTS with dist.GradBucket / Future[Tensor] (distributed.py:2268-2289); postponed (string) annotations fail it.
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .comm import P2PComm, TensorTable, dtype_code


class AllreduceState:
    """The hook's ``state``: one ``P2PComm`` (its arena is the staging the peers read), a side stream, and the
    chunking of buckets larger than the staging area (iteration 0 of torch >= 2.x reduces ALL gradients as one
    bucket, distributed.py:1224-1225: ~100 MB for ResNet-50).  Construction is collective over the default
    process group."""

    def __init__(self, device: Optional[int] = None, staging_mb: int = 64, comm: Optional[P2PComm] = None):
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        #: two staging regions, used alternately, so a chunk's flatten may run while the previous chunk's results
        #: are still being copied out
        self.region = int(staging_mb) << 20
        self.comm = comm if comm is not None else P2PComm.from_process_group(2 * self.region + (1 << 20), device)
        self.world = self.comm.world
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.max_ctas = int(os.environ.get("B200DDP_CTAS", "0"))
        self._turn = 0
        self.buckets_reduced = 0

    def close(self):
        if self.comm is not None:
            torch.cuda.synchronize(self.device)
            self.comm.destroy()
            self.comm = None


def allreduce_hook(state: AllreduceState, bucket: dist.GradBucket) -> torch.futures.Future[torch.Tensor]:
    """flatten (one tensor: torch's flat bucket) -> allreduce with the 1/world scale folded in -> copy back, all on
    the state's side stream; the returned future completes when that stream has (CUDA-aware future: consumers'
    streams wait on an event, the host does not block)."""
    buf = bucket.buffer()
    comm, s = state.comm, state.stream.cuda_stream
    code = dtype_code(buf.dtype)
    es = buf.element_size()
    total = buf.numel()
    per = state.region // es // (16 // es) * (16 // es)  # elements per chunk, a whole number of 16-byte vectors
    comm.stream_wait(s, torch.cuda.current_stream(state.device).cuda_stream)
    base = buf.data_ptr()
    done = 0
    while done < total:
        n = min(per, total - done)
        off = (state._turn & 1) * state.region
        state._turn += 1
        table = TensorTable([base + done * es], [0], [n])
        nbytes = (n * es + 15) // 16 * 16
        comm.flatten_scale(table, code, off, 1.0, s)
        comm.allreduce(off, off, nbytes, code, _lib.ALGO_AUTO, state.max_ctas, s, scale=1.0 / state.world)
        comm.unflatten(table, code, off, s)
        done += n
    state.buckets_reduced += 1
    comm.poll_error()
    fut: torch.futures.Future[torch.Tensor] = torch.futures.Future(devices=[state.device])
    with torch.cuda.stream(state.stream):
        fut.set_result(buf)
    return fut
