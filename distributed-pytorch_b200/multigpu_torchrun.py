"""``multigpu_torchrun.py`` surface (ref:multigpu_torchrun.py) bound to the B200-native path: torchrun
launch (rank/world from the environment), snapshot + resume.

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node N \\
        distributed-pytorch_b200/multigpu_torchrun.py <total_epochs> <save_every>
"""
from __future__ import annotations

import os
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import b200ddp  # noqa: F401
    __package__ = "b200ddp"

import torch

from . import DistributedDataParallel as DDP
from . import destroy_process_group, init_process_group
from ._trainer_core import TrainerCore
from ._trainer_core import sharded_loader as prepare_dataloader
from ._trainer_core import toy_train_objs as load_train_objs

__all__ = ["ddp_setup", "Trainer", "load_train_objs", "prepare_dataloader", "main"]


def ddp_setup() -> None:
    """RANK / WORLD_SIZE / MASTER_* come from torchrun (ref:multigpu_torchrun.py:12-13)."""
    init_process_group(backend="nccl")


class Trainer(TrainerCore):
    def __init__(self, model: torch.nn.Module, train_data, optimizer: torch.optim.Optimizer, save_every: int,
                 snapshot_path: str) -> None:
        local = int(os.environ["LOCAL_RANK"])
        self.epochs_run = 0
        self.snapshot_path = snapshot_path
        self.model = model.to(local)
        if os.path.exists(snapshot_path):
            # resume into the bare model BEFORE it is wrapped (ref:multigpu_torchrun.py:30-34)
            print("Loading snapshot")
            self._load_snapshot(snapshot_path)
        self._adopt(DDP(self.model, device_ids=[local]), train_data, optimizer, local, save_every)

    def _load_snapshot(self, snapshot_path: str) -> None:
        state = torch.load(snapshot_path, map_location=f"cuda:{int(os.environ['LOCAL_RANK'])}")
        self.model.load_state_dict(state["MODEL_STATE"])
        self.epochs_run = state["EPOCHS_RUN"]
        print(f"Resuming training from snapshot at Epoch {self.epochs_run}")

    def _save_snapshot(self, epoch: int) -> None:
        self._check_health()  # never save weights that took a step on a timed-out (un-averaged) bucket
        torch.save({"EPOCHS_RUN": epoch, "MODEL_STATE": self.model.module.state_dict()}, self.snapshot_path)
        print(f"Epoch {epoch} | Training snapshot saved at {self.snapshot_path}")

    def train(self, max_epochs: int) -> None:
        for epoch in range(self.epochs_run, max_epochs):
            self._run_epoch(epoch)
            if epoch % self.save_every == 0 and self.gpu_id == 0:
                self._save_snapshot(epoch)


def main(save_every: int, total_epochs: int, snapshot_path: str = "snapshot.pt") -> None:
    ddp_setup()
    train_set, net, sgd = load_train_objs()
    Trainer(net, prepare_dataloader(train_set, batch_size=32), sgd, save_every, snapshot_path).train(total_epochs)
    destroy_process_group()


if __name__ == "__main__":
    main(total_epochs=int(sys.argv[1]), save_every=int(sys.argv[2]))
