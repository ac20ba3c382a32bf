"""``multigpu.py`` surface (ref:multigpu.py): ``ddp_setup``, ``Trainer``, ``load_train_objs``,
``prepare_dataloader``, ``main`` with the reference's signatures, where ``DDP`` is
``b200ddp.DistributedDataParallel`` and the process group comes from ``b200ddp.init_process_group``.

Run:  python -m b200ddp.multigpu <total_epochs> <save_every>
"""
from __future__ import annotations

import os
import sys

if __package__ in (None, ""):  # executed as a file: make the alias package importable
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import b200ddp  # noqa: F401
    __package__ = "b200ddp"

import torch
import torch.multiprocessing as mp
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from . import DistributedDataParallel as DDP
from . import destroy_process_group, init_process_group
from ._trainer_core import TrainerCore
from .utils import MyTrainDataset

CHECKPOINT_PATH = "checkpoint.pt"


def ddp_setup(rank: int, world_size: int) -> None:
    """rank: this process's id; world_size: number of processes (ref:multigpu.py:12-20)."""
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12355")
    init_process_group(backend="nccl", rank=rank, world_size=world_size)


class Trainer(TrainerCore):
    def __init__(self, model: torch.nn.Module, train_data: DataLoader, optimizer: torch.optim.Optimizer,
                 gpu_id: int, save_every: int) -> None:
        self.gpu_id = gpu_id
        self.train_data = train_data
        self.optimizer = optimizer
        self.save_every = save_every
        self.model = DDP(model.to(gpu_id), device_ids=[gpu_id])

    def _save_checkpoint(self, epoch: int) -> None:
        torch.save(self.model.module.state_dict(), CHECKPOINT_PATH)
        print(f"Epoch {epoch} | Training checkpoint saved at {CHECKPOINT_PATH}")

    def train(self, max_epochs: int) -> None:
        for epoch in range(max_epochs):
            self._run_epoch(epoch)
            if self.gpu_id == 0 and epoch % self.save_every == 0:
                self._save_checkpoint(epoch)


def load_train_objs():
    train_set = MyTrainDataset(2048)
    model = torch.nn.Linear(20, 1)
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-3)
    return train_set, model, optimizer


def prepare_dataloader(dataset: Dataset, batch_size: int) -> DataLoader:
    return DataLoader(dataset, batch_size=batch_size, pin_memory=True, shuffle=False,
                      sampler=DistributedSampler(dataset))


def main(rank: int, world_size: int, save_every: int, total_epochs: int) -> None:
    ddp_setup(rank, world_size)
    dataset, model, optimizer = load_train_objs()
    train_data = prepare_dataloader(dataset, batch_size=32)
    Trainer(model, train_data, optimizer, rank, save_every).train(total_epochs)
    destroy_process_group()


if __name__ == "__main__":
    epochs, every = int(sys.argv[1]), int(sys.argv[2])
    n = torch.cuda.device_count()
    # the reference passes (world_size, total_epochs, save_every) here, which swaps the last two
    # (SURVEY §9 Q3); the arguments are passed in main()'s order instead
    mp.spawn(main, args=(n, every, epochs), nprocs=n)
