"""``multigpu_profile.py`` surface (ref:multigpu_profile.py): ResNet-50 on synthetic 3x224x224 with the
same ``Trainer(model, train_set, optimizer, device, batch_size, profile)`` / ``run_batch`` / ``run_epoch`` /
``save_checkpoint`` / ``ddp_setup`` / ``main`` names — the torch.profiler/tensorboard capture
(ref:multigpu_profile.py:80-91) replaced by what profiles OUR kernels: a CUDA-profiler window
(``cudaProfilerStart/Stop``) with the reference's wait=1 / warmup=1 / active=5 schedule plus one NVTX range
per step, so that

    ncu --profile-from-start off --set full --clock-control none -k regex:'twoshot|nvls|multi_tensor|fused' \\
        -o gpurun_out/prof python -m b200ddp.multigpu_profile        # ONE GPU (ncu replays kernels)

captures exactly the active steps.  Run without ncu it is a plain 3-epoch training run.
"""
from __future__ import annotations

import os
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import b200ddp  # noqa: F401
    __package__ = "b200ddp"

import torch
import torch.multiprocessing as mp
import torch.nn.functional as F
from torch.utils.data.distributed import DistributedSampler

from . import DistributedDataParallel as DDP
from . import destroy_process_group, init_process_group
from .utils import MyRandomDataset


def load_train_objs():
    from torchvision.models import resnet50

    train_set = MyRandomDataset(2048, (3, 224, 224))
    model = resnet50()
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-3)
    return train_set, model, optimizer


class CudaProfilerWindow:
    """start()/step()/stop() like torch.profiler.profile with schedule(wait, warmup, active): opens the CUDA
    profiler range for the `active` steps only and wraps every step in an NVTX range."""

    def __init__(self, wait: int = 1, warmup: int = 1, active: int = 5) -> None:
        self.first = wait + warmup
        self.last = wait + warmup + active
        self.n = 0
        self.open = False
        self.in_range = False

    def _push(self) -> None:
        torch.cuda.nvtx.range_push(f"b200ddp step {self.n}")
        self.in_range = True

    def start(self) -> None:
        self.n = 0
        self._push()

    def step(self) -> None:
        if self.in_range:
            torch.cuda.nvtx.range_pop()
            self.in_range = False
        self.n += 1
        if self.n == self.first and not self.open:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            self.open = True
        elif self.n == self.last and self.open:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            self.open = False
        self._push()

    def stop(self) -> None:
        if self.in_range:
            torch.cuda.nvtx.range_pop()
            self.in_range = False
        if self.open:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            self.open = False


class Trainer:
    def __init__(self, model, train_set, optimizer, device, batch_size, profile: bool = False) -> None:
        self.model = self.prepare_model_for_DDP(model, device)
        self.train_set = train_set
        self.optimizer = optimizer
        self.device = device
        self.batch_size = batch_size
        self.profiler = self._create_profiler() if profile else None

    def prepare_model_for_DDP(self, model, device):
        model.to(device)
        return DDP(model, [device])

    def run_batch(self, source, targets):
        self.optimizer.zero_grad()
        source, targets = source.to(self.device), targets.to(self.device)
        loss = F.cross_entropy(self.model(source), targets)
        loss.backward()
        self.optimizer.step()

    def run_epoch(self, nb_epochs: int) -> None:
        train_data = torch.utils.data.DataLoader(self.train_set, batch_size=32, shuffle=False,
                                                 sampler=DistributedSampler(self.train_set))
        if self.profiler:
            self.profiler.start()
        for epoch in range(nb_epochs):
            print(f"[GPU{self.device}] Epoch {epoch} | No. of batches: {len(train_data)} | "
                  f"Batchsize: {len(next(iter(train_data))[0])}")
            for source, targets in train_data:
                self.run_batch(source, targets)
                if self.profiler:
                    self.profiler.step()
        if self.profiler:
            self.profiler.stop()

    def save_checkpoint(self) -> None:
        torch.save(self.model.state_dict(), "model_ddp.pth")  # wrapped state dict: module.-prefixed keys
        print("Model state dict saved at model_ddp.pth")

    def _create_profiler(self) -> CudaProfilerWindow:
        return CudaProfilerWindow(wait=1, warmup=1, active=5)


def ddp_setup(rank: int, world_size: int) -> None:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12355")
    init_process_group("nccl", rank=rank, world_size=world_size)


def main(rank: int, world_size: int) -> None:
    ddp_setup(rank, world_size)
    dataset, model, optimizer = load_train_objs()
    trainer = Trainer(model, dataset, optimizer, device=rank, batch_size=32, profile=True)
    trainer.run_epoch(3)
    if rank == 0:
        trainer.save_checkpoint()
    destroy_process_group()


if __name__ == "__main__":
    n = torch.cuda.device_count()
    mp.spawn(main, args=(n,), nprocs=n)
