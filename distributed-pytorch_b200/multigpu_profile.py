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
    """ResNet-50 (random init), 2048 synthetic 3x224x224 samples with 1000-way soft targets, SGD lr 1e-3
    (ref:multigpu_profile.py:13-27; the commented-out vit_l_32 alternative is not carried over)."""
    from torchvision.models import resnet50

    net = resnet50()
    return MyRandomDataset(2048, (3, 224, 224)), net, torch.optim.SGD(net.parameters(), lr=1e-3)


class CudaProfilerWindow:
    """start()/step()/stop() with torch.profiler's schedule semantics (wait, warmup, active): the CUDA
    profiler range (cudaProfilerStart/Stop, honoured by ``ncu --profile-from-start off``) is open for the
    ``active`` steps only, and every step sits inside its own NVTX range."""

    def __init__(self, wait: int = 1, warmup: int = 1, active: int = 5) -> None:
        self.first, self.last = wait + warmup, wait + warmup + active
        self.n, self.open, self.in_range = 0, False, False

    def _nvtx(self, push: bool) -> None:
        if push:
            torch.cuda.nvtx.range_push(f"b200ddp step {self.n}")
        elif self.in_range:
            torch.cuda.nvtx.range_pop()
        self.in_range = push

    def _window(self, open_it: bool) -> None:
        torch.cuda.synchronize()
        (torch.cuda.profiler.start if open_it else torch.cuda.profiler.stop)()
        self.open = open_it

    def start(self) -> None:
        self.n = 0
        self._nvtx(True)

    def step(self) -> None:
        self._nvtx(False)
        self.n += 1
        if self.n == self.first and not self.open:
            self._window(True)
        elif self.n == self.last and self.open:
            self._window(False)
        self._nvtx(True)

    def stop(self) -> None:
        self._nvtx(False)
        if self.open:
            self._window(False)


class Trainer:
    """Same constructor and method names as ref:multigpu_profile.py:30-91."""

    def __init__(self, model, train_set, optimizer, device, batch_size, profile: bool = False) -> None:
        self.device, self.batch_size = device, batch_size
        self.train_set, self.optimizer = train_set, optimizer
        self.model = self.prepare_model_for_DDP(model, device)
        self.profiler = self._create_profiler() if profile else None

    def prepare_model_for_DDP(self, model, device):
        return DDP(model.to(device), [device])

    def run_batch(self, source, targets) -> None:
        self.optimizer.zero_grad()
        logits = self.model(source.to(self.device))
        F.cross_entropy(logits, targets.to(self.device)).backward()  # soft-label CE, ref:multigpu_profile.py:49
        self.optimizer.step()

    def _loader(self):
        # the reference hard-codes 32 here and ignores its batch_size argument; kept
        return torch.utils.data.DataLoader(self.train_set, sampler=DistributedSampler(self.train_set), batch_size=32,
                                           shuffle=False)

    def run_epoch(self, nb_epochs: int) -> None:
        batches = self._loader()
        prof = self.profiler
        if prof:
            prof.start()
        for epoch in range(nb_epochs):
            width = len(next(iter(batches))[0])
            print(f"[GPU{self.device}] Epoch {epoch} | No. of batches: {len(batches)} | Batchsize: {width}")
            for source, targets in batches:
                self.run_batch(source, targets)
                if prof:
                    prof.step()
        if prof:
            prof.stop()

    def save_checkpoint(self, path: str = "model_ddp.pth") -> None:
        self.model.check_health()  # never save weights that took a step on a timed-out (un-averaged) bucket
        torch.save(self.model.state_dict(), path)  # the WRAPPED state dict: keys carry the module. prefix
        print(f"Model state dict saved at {path}")

    def _create_profiler(self) -> CudaProfilerWindow:
        return CudaProfilerWindow(wait=1, warmup=1, active=5)  # the reference's schedule, ref:multigpu_profile.py:82


def ddp_setup(rank: int, world_size: int) -> None:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12355")
    init_process_group("nccl", rank=rank, world_size=world_size)


def main(rank: int, world_size: int) -> None:
    ddp_setup(rank, world_size)
    train_set, net, sgd = load_train_objs()
    trainer = Trainer(net, train_set, sgd, device=rank, batch_size=32, profile=True)
    trainer.run_epoch(3)
    if rank == 0:
        trainer.save_checkpoint()
    destroy_process_group()


if __name__ == "__main__":
    gpus = torch.cuda.device_count()
    mp.spawn(main, nprocs=gpus, args=(gpus,))
