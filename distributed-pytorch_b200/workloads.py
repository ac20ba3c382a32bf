"""The BASELINE.json workloads: model builders, synthetic batches and losses shared by bench.py, the
tests and the profiling script.  Random-init weights and synthetic data only (no network, no datasets).

  toy       ``Linear(20, 1)`` on ``MyTrainDataset``-shaped rows (ref:multigpu.py:65-69)          configs[0], [1]
  wide      ``Linear(4096, 4096)`` — "toy model widened to 4096-dim", one 64 MiB bucket           configs[2]
  slices    ``k`` x ``Linear(4096, 64)`` heads: ~1 MiB parameters each, so ``bucket_cap_mb`` sweeps
            produce buckets of every size (SURVEY §8 d)                                           configs[2]
  resnet50  torchvision ``resnet50()`` on 3x224x224 with soft-label CE (ref:multigpu_profile.py)  configs[3]
  gpt2m     24 x 1024 pre-LN transformer, vocab 50257, seq 1024, bf16                             configs[4]
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class SliceHeads(nn.Module):
    """``k`` independent ``Linear(dim, width)`` heads on the same input, summed into one output column
    block.  Each head is ~dim*width*4 bytes, so gradient buckets of any size can be formed."""

    def __init__(self, k: int = 64, dim: int = 4096, width: int = 64) -> None:
        super().__init__()
        self.heads = nn.ModuleList(nn.Linear(dim, width) for _ in range(k))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.heads[0](x)
        for h in self.heads[1:]:
            out = out + h(x)
        return out


class _Block(nn.Module):
    def __init__(self, d: int, heads: int, ffn: int) -> None:
        super().__init__()
        self.ln1 = nn.LayerNorm(d)
        self.qkv = nn.Linear(d, 3 * d)
        self.proj = nn.Linear(d, d)
        self.ln2 = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.heads = heads

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, t, d = x.shape
        q, k, v = self.qkv(self.ln1(x)).view(b, t, 3, self.heads, d // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b, t, d)
        x = x + self.proj(a)
        return x + self.fc2(F.gelu(self.fc1(self.ln2(x))))


class GPT2Shaped(nn.Module):
    """GPT-2-medium-shaped decoder: 24 layers, d=1024, 16 heads, ffn 4096, vocab 50257, ctx 1024; the LM
    head is tied to the token embedding as in GPT-2 (354.8 M parameters, SURVEY §8 d)."""

    def __init__(self, layers: int = 24, d: int = 1024, heads: int = 16, ffn: int = 4096, vocab: int = 50257,
                 ctx: int = 1024) -> None:
        super().__init__()
        self.tok = nn.Embedding(vocab, d)
        self.pos = nn.Embedding(ctx, d)
        self.blocks = nn.ModuleList(_Block(d, heads, ffn) for _ in range(layers))
        self.ln_f = nn.LayerNorm(d)

    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        t = idx.shape[1]
        x = self.tok(idx) + self.pos(torch.arange(t, device=idx.device))
        for blk in self.blocks:
            x = blk(x)
        return F.linear(self.ln_f(x), self.tok.weight)


@dataclass
class Workload:
    name: str
    build: Callable[[], nn.Module]
    batch: Callable[[int, torch.device, Optional[torch.Generator]], Tuple[torch.Tensor, torch.Tensor]]
    loss: Callable[[torch.Tensor, torch.Tensor], torch.Tensor]
    dtype: torch.dtype = torch.float32
    per_gpu_batch: int = 32  # ref:multigpu.py:85


def _rand(shape, device, gen, dtype=torch.float32):
    # generated on the host with a CPU generator so that a (seed, rank) pair means the same rows everywhere
    return torch.rand(shape, generator=gen, dtype=torch.float32).to(device=device, dtype=dtype)


def _rows(features: int, targets: int):
    def make(b: int, device, gen=None):
        return _rand((b, features), device, gen), _rand((b, targets), device, gen)
    return make


def _images(b: int, device, gen=None):
    return _rand((b, 3, 224, 224), device, gen), _rand((b, 1000), device, gen)


def _tokens(b: int, device, gen=None):
    idx = torch.randint(0, 50257, (b, 1024), generator=gen).to(device)
    return idx, idx


def _lm_loss(logits: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), idx[:, 1:].reshape(-1))


def _resnet50() -> nn.Module:
    from torchvision.models import resnet50

    return resnet50()


WORKLOADS = {
    # the reference's toy objects (ref:multigpu.py:65-69); CE on a single logit gives all-zero gradients
    # (SURVEY §9 Q1), so the measured variant uses the multinode script's MSELoss (ref:multinode_torchrun.py:46)
    "toy": Workload("toy", lambda: nn.Linear(20, 1), _rows(20, 1), F.mse_loss),
    "toy_ce": Workload("toy_ce", lambda: nn.Linear(20, 1), _rows(20, 1), nn.CrossEntropyLoss()),
    "wide": Workload("wide", lambda: nn.Linear(4096, 4096), _rows(4096, 4096), F.mse_loss),
    "slices": Workload("slices", lambda: SliceHeads(64, 4096, 64), _rows(4096, 64), F.mse_loss),
    "resnet50": Workload("resnet50", _resnet50, _images, F.cross_entropy),
    "gpt2m": Workload("gpt2m", lambda: GPT2Shaped().to(torch.bfloat16), _tokens, _lm_loss, torch.bfloat16, 8),
    "mlp_bf16": Workload("mlp_bf16", lambda: nn.Sequential(nn.Linear(20, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                                                           nn.Linear(256, 7)).to(torch.bfloat16),
                         lambda b, device, gen=None: tuple(t.to(torch.bfloat16) for t in _rows(20, 7)(b, device, gen)),
                         F.mse_loss, torch.bfloat16),
    # the GPT-2-medium shape cut to 2 layers for parity tests: same tensor shapes and bucket structure per layer,
    # the same 103 MB tied-embedding bucket, a fraction of the time
    "gpt2m_2l": Workload("gpt2m_2l", lambda: GPT2Shaped(layers=2).to(torch.bfloat16),
                         lambda b, device, gen=None: tuple(t[:, :128] for t in _tokens(b, device, gen)), _lm_loss,
                         torch.bfloat16, 8),
    # small multi-bucket model for parity tests
    "mlp": Workload("mlp", lambda: nn.Sequential(nn.Linear(20, 96), nn.ReLU(), nn.Linear(96, 96), nn.ReLU(),
                                                 nn.Linear(96, 7)), _rows(20, 7), F.mse_loss),
}


def get(name: str) -> Workload:
    try:
        return WORKLOADS[name]
    except KeyError:
        raise ValueError(f"unknown workload {name!r}; choose from {sorted(WORKLOADS)}") from None
