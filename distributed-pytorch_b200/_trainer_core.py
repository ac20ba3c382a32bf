"""Shared body of the ``Trainer`` classes the surface scripts expose.

One optimisation step is the reference's ``_run_batch`` (ref:multigpu.py:38-43): zero_grad -> forward
-> loss -> backward -> step.  The gradient allreduce fires inside ``loss.backward()`` through the hooks
``b200ddp.DistributedDataParallel`` installed — nothing here knows about it, exactly as in the reference.

Opt-in extension (not in the reference): ``enable_cuda_graph`` captures that same step — including the
bucket kernels our Reducer enqueues on its side stream — into one CUDA graph, so a launch-bound step (the
toy model's ~15 tiny kernels) is replayed with a single launch.  This works because the allreduce kernels
keep their barrier generation in device memory (no per-launch host state, see csrc/kernels.cu).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler


def toy_train_objs() -> Tuple[Dataset, torch.nn.Module, torch.optim.Optimizer]:
    """The objects every toy script trains (ref:multigpu.py:65-69, ref:multigpu_torchrun.py:71-75): 2048
    ``MyTrainDataset`` rows, ``Linear(20, 1)``, SGD with lr 1e-3 built on the UN-wrapped model."""
    from .utils import MyTrainDataset

    net = torch.nn.Linear(MyTrainDataset.FEATURES, MyTrainDataset.TARGETS)
    return MyTrainDataset(2048), net, torch.optim.SGD(net.parameters(), lr=1e-3)


def sharded_loader(dataset: Dataset, batch_size: int) -> DataLoader:
    """Per-rank loader (ref:multigpu.py:72-79): pinned memory, no loader-level shuffle, rows chosen by a
    default ``DistributedSampler`` (which shuffles with seed 0 and is never re-seeded)."""
    return DataLoader(dataset, batch_size=batch_size, sampler=DistributedSampler(dataset), shuffle=False,
                      pin_memory=True)


class TrainerCore:
    #: loss used by ``_run_batch``.  The reference's choice (ref:multigpu.py:41) — note SURVEY §9 Q1: with
    #: the toy Linear(20,1) head this loss, and therefore every gradient, is identically zero.
    loss_fn: Callable = staticmethod(torch.nn.CrossEntropyLoss())

    model: torch.nn.Module
    optimizer: torch.optim.Optimizer
    train_data: DataLoader
    gpu_id: int

    _graph: Optional[torch.cuda.CUDAGraph] = None
    _g_source: Optional[torch.Tensor] = None
    _g_targets: Optional[torch.Tensor] = None
    _g_loss: Optional[torch.Tensor] = None
    graph_kernel_launches: int = 0  # b200ddp kernels captured per replayed step

    def _adopt(self, wrapped: torch.nn.Module, train_data, optimizer: torch.optim.Optimizer, gpu_id: int,
               save_every: int) -> None:
        """Common constructor tail of the script Trainers."""
        self.gpu_id, self.save_every = gpu_id, save_every
        self.train_data, self.optimizer = train_data, optimizer
        self.model = wrapped

    def _eager_step(self, source: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        self.optimizer.zero_grad()
        loss = self.loss_fn(self.model(source), targets)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _run_batch(self, source: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if self._graph is None:
            return self._eager_step(source, targets)
        self._g_source.copy_(source, non_blocking=True)
        self._g_targets.copy_(targets, non_blocking=True)
        self._graph.replay()
        return self._g_loss

    def enable_cuda_graph(self, source: torch.Tensor, targets: torch.Tensor, warmup: int = 3) -> None:
        """Capture one full step on batches shaped like (source, targets); later ``_run_batch`` calls copy
        their batch into the captured input buffers and replay.  The ``warmup`` eager steps are real
        training steps (they also trigger the Reducer's one-time bucket rebuild before capture).  All
        ranks must call this collectively: the captured allreduce kernels synchronise across ranks."""
        if self._graph is not None:
            return
        dev = source.device
        self._g_source, self._g_targets = source.clone(), targets.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):
                self._eager_step(self._g_source, self._g_targets)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        comm = getattr(getattr(self.model, "engine", None), "comm", None)
        n0 = comm.launch_count if comm is not None else 0
        self.optimizer.zero_grad(set_to_none=True)  # gradients are (re)created inside the graph's pool
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self.loss_fn(self.model(self._g_source), self._g_targets)
            loss.backward()
            self.optimizer.step()
            self._g_loss = loss.detach()
        self.graph_kernel_launches = (comm.launch_count - n0) if comm is not None else 0
        self._graph = graph

    def _run_epoch(self, epoch: int) -> None:
        steps = len(self.train_data)
        first = next(iter(self.train_data))[0]
        print(f"[GPU{self.gpu_id}] Epoch {epoch} | Batchsize: {len(first)} | Steps: {steps}")
        for source, targets in self.train_data:
            if self._graph is not None:  # the replayed step copies host batches straight into its input buffers
                self._run_batch(source, targets)
            else:
                self._run_batch(source.to(self.gpu_id), targets.to(self.gpu_id))
