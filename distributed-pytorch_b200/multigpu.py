"""``multigpu.py`` surface (ref:multigpu.py) bound to the B200-native path.

Exports the reference script's names with the reference's signatures — ``ddp_setup(rank, world_size)``,
``Trainer(model, train_data, optimizer, gpu_id, save_every)`` with ``_run_batch`` / ``_run_epoch`` /
``_save_checkpoint`` / ``train``, ``load_train_objs()``, ``prepare_dataloader(dataset, batch_size)``,
``main(rank, world_size, save_every, total_epochs)`` — where ``DDP`` is ``b200ddp.DistributedDataParallel``
and the process group comes from ``b200ddp.init_process_group``.  The step itself lives in
``_trainer_core.TrainerCore`` (shared with the torchrun variant).

    python -m b200ddp.multigpu <total_epochs> <save_every>
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

from . import DistributedDataParallel as DDP
from . import destroy_process_group, init_process_group
from ._trainer_core import TrainerCore
from ._trainer_core import sharded_loader as prepare_dataloader  # (dataset, batch_size) -> DataLoader
from ._trainer_core import toy_train_objs as load_train_objs  # () -> (train_set, model, optimizer)

CHECKPOINT_PATH = "checkpoint.pt"
RENDEZVOUS = {"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "12355"}  # ref:multigpu.py:18-19 uses localhost:12355


def ddp_setup(rank: int, world_size: int) -> None:
    """``rank``: this process; ``world_size``: number of processes (one per GPU)."""
    for key, value in RENDEZVOUS.items():
        os.environ.setdefault(key, value)
    init_process_group(backend="nccl", rank=rank, world_size=world_size)


class Trainer(TrainerCore):
    def __init__(self, model: torch.nn.Module, train_data, optimizer: torch.optim.Optimizer, gpu_id: int,
                 save_every: int) -> None:
        # ref:multigpu.py:32,36 — move to the GPU, then wrap; the optimizer already holds the parameters
        self._adopt(DDP(model.to(gpu_id), device_ids=[gpu_id]), train_data, optimizer, gpu_id, save_every)

    def _save_checkpoint(self, epoch: int) -> None:
        """Rank 0 writes the UN-wrapped weights (``.module``), like ref:multigpu.py:53-56."""
        self._check_health()  # never save weights that took a step on a timed-out (un-averaged) bucket
        torch.save(self.model.module.state_dict(), CHECKPOINT_PATH)
        print(f"Epoch {epoch} | Training checkpoint saved at {CHECKPOINT_PATH}")

    def train(self, max_epochs: int) -> None:
        for epoch in range(max_epochs):
            self._run_epoch(epoch)
            if epoch % self.save_every == 0 and self.gpu_id == 0:
                self._save_checkpoint(epoch)


def main(rank: int, world_size: int, save_every: int, total_epochs: int) -> None:
    ddp_setup(rank, world_size)
    train_set, net, sgd = load_train_objs()
    Trainer(net, prepare_dataloader(train_set, batch_size=32), sgd, rank, save_every).train(total_epochs)
    destroy_process_group()


if __name__ == "__main__":
    epochs, every = (int(a) for a in sys.argv[1:3])
    gpus = torch.cuda.device_count()
    # the reference hands mp.spawn (world_size, total_epochs, save_every), which swaps main()'s last two
    # parameters; here they are passed in main()'s own order
    mp.spawn(main, nprocs=gpus, args=(gpus, every, epochs))
