"""Synthetic datasets with the reference's names and item shapes (ref:utils.py:4-26).

``MyTrainDataset(size)`` yields ``(x[20], y[1])`` float32 pairs drawn from U[0,1) once at construction
(ref:utils.py:7); ``MyRandomDataset(size, input_shape, target_shape)`` draws a fresh pair per access
(ref:utils.py:26).  Rows are stored as two dense tensors rather than a Python list of tuples, which keeps
the item interface and lets a whole shard be moved to the GPU in one copy (``device_shard``).  An optional
``seed`` makes rows reproducible for the parity harness; the reference never seeds.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch.utils.data import Dataset


def _generator(seed: Optional[int]) -> Optional[torch.Generator]:
    if seed is None:
        return None
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


class MyTrainDataset(Dataset):
    FEATURES = 20
    TARGETS = 1

    def __init__(self, size: int, seed: Optional[int] = None, features: int = FEATURES, targets: int = TARGETS):
        self.size = int(size)
        g = _generator(seed)
        self.x = torch.rand(self.size, features, generator=g)
        self.y = torch.rand(self.size, targets, generator=g)

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, index: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.x[index], self.y[index]

    def device_shard(self, indices: Sequence[int], device) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rows ``indices`` (a rank's DistributedSampler shard) as two device-resident tensors."""
        idx = torch.as_tensor(list(indices), dtype=torch.long)
        return self.x[idx].to(device), self.y[idx].to(device)


class DeviceShardLoader:
    """A rank's ``DistributedSampler`` shard of a ``MyTrainDataset`` kept RESIDENT on the device and served
    as batch views — the same rows, in the same order and with the same batch boundaries as
    ``prepare_dataloader(dataset, batch_size)`` (ref:multigpu.py:72-79: ``shuffle=False`` on the loader,
    ``DistributedSampler`` defaults, last batch kept), but with one H2D copy per epoch-set instead of one per
    batch and no per-sample collation.  Iterating yields ``(source, targets)`` device tensors."""

    def __init__(self, dataset: "MyTrainDataset", batch_size: int, device, sampler=None) -> None:
        from torch.utils.data.distributed import DistributedSampler

        self.sampler = sampler if sampler is not None else DistributedSampler(dataset)
        self.batch_size = int(batch_size)
        self.x, self.y = dataset.device_shard(list(self.sampler), device)

    def __len__(self) -> int:
        return (self.x.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, self.x.shape[0], self.batch_size):
            yield self.x[i:i + self.batch_size], self.y[i:i + self.batch_size]


class MyRandomDataset(Dataset):
    def __init__(self, size: int, input_shape: Sequence[int], target_shape: Sequence[int] = (1000,),
                 seed: Optional[int] = None):
        self.size = int(size)
        self.shape = tuple(input_shape)
        self.target_shape = tuple(target_shape)
        self._gen = _generator(seed)

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, index: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return torch.rand(self.shape, generator=self._gen), torch.rand(self.target_shape, generator=self._gen)
