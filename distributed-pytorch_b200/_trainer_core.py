"""Shared body of the ``Trainer`` classes the surface scripts expose.

One optimisation step is the reference's ``_run_batch`` (ref:multigpu.py:38-43): zero_grad -> forward
-> loss -> backward -> step.  The gradient allreduce fires inside ``loss.backward()`` through the hooks
``b200ddp.DistributedDataParallel`` installed — nothing here knows about it, exactly as in the reference.
"""
from __future__ import annotations

from typing import Callable

import torch
from torch.utils.data import DataLoader


class TrainerCore:
    #: loss used by ``_run_batch``.  The reference's choice (ref:multigpu.py:41) — note SURVEY §9 Q1: with
    #: the toy Linear(20,1) head this loss, and therefore every gradient, is identically zero.
    loss_fn: Callable = staticmethod(torch.nn.CrossEntropyLoss())

    model: torch.nn.Module
    optimizer: torch.optim.Optimizer
    train_data: DataLoader
    gpu_id: int

    def _run_batch(self, source: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        self.optimizer.zero_grad()
        loss = self.loss_fn(self.model(source), targets)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _run_epoch(self, epoch: int) -> None:
        steps = len(self.train_data)
        first = next(iter(self.train_data))[0]
        print(f"[GPU{self.gpu_id}] Epoch {epoch} | Batchsize: {len(first)} | Steps: {steps}")
        for source, targets in self.train_data:
            self._run_batch(source.to(self.gpu_id), targets.to(self.gpu_id))
