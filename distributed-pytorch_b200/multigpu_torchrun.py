"""``multigpu_torchrun.py`` surface (ref:multigpu_torchrun.py): torchrun launch, snapshot/resume.

Run:  torchrun --standalone --nproc-per-node N distributed-pytorch_b200/multigpu_torchrun.py <total_epochs> <save_every>
"""
from __future__ import annotations

import os
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import b200ddp  # noqa: F401
    __package__ = "b200ddp"

import torch
from torch.utils.data import DataLoader

from . import DistributedDataParallel as DDP
from . import destroy_process_group, init_process_group
from ._trainer_core import TrainerCore
from .multigpu import load_train_objs, prepare_dataloader  # same objects as ref:multigpu_torchrun.py:71-86

__all__ = ["ddp_setup", "Trainer", "load_train_objs", "prepare_dataloader", "main"]


def ddp_setup() -> None:
    """torchrun provides RANK / WORLD_SIZE / MASTER_* (ref:multigpu_torchrun.py:12-13)."""
    init_process_group(backend="nccl")


class Trainer(TrainerCore):
    def __init__(self, model: torch.nn.Module, train_data: DataLoader, optimizer: torch.optim.Optimizer,
                 save_every: int, snapshot_path: str) -> None:
        self.gpu_id = int(os.environ["LOCAL_RANK"])
        self.model = model.to(self.gpu_id)
        self.train_data = train_data
        self.optimizer = optimizer
        self.save_every = save_every
        self.epochs_run = 0
        self.snapshot_path = snapshot_path
        if os.path.exists(snapshot_path):  # resume BEFORE wrapping (ref:multigpu_torchrun.py:30-34)
            print("Loading snapshot")
            self._load_snapshot(snapshot_path)
        self.model = DDP(self.model, device_ids=[self.gpu_id])

    def _load_snapshot(self, snapshot_path: str) -> None:
        snapshot = torch.load(snapshot_path, map_location=f"cuda:{self.gpu_id}")
        self.model.load_state_dict(snapshot["MODEL_STATE"])
        self.epochs_run = snapshot["EPOCHS_RUN"]
        print(f"Resuming training from snapshot at Epoch {self.epochs_run}")

    def _save_snapshot(self, epoch: int) -> None:
        torch.save({"MODEL_STATE": self.model.module.state_dict(), "EPOCHS_RUN": epoch}, self.snapshot_path)
        print(f"Epoch {epoch} | Training snapshot saved at {self.snapshot_path}")

    def train(self, max_epochs: int) -> None:
        for epoch in range(self.epochs_run, max_epochs):
            self._run_epoch(epoch)
            if self.gpu_id == 0 and epoch % self.save_every == 0:
                self._save_snapshot(epoch)


def main(save_every: int, total_epochs: int, snapshot_path: str = "snapshot.pt") -> None:
    ddp_setup()
    dataset, model, optimizer = load_train_objs()
    train_data = prepare_dataloader(dataset, batch_size=32)
    Trainer(model, train_data, optimizer, save_every, snapshot_path).train(total_epochs)
    destroy_process_group()


if __name__ == "__main__":
    main(save_every=int(sys.argv[2]), total_epochs=int(sys.argv[1]))
