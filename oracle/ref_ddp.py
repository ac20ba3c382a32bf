"""The reference stack's OWN implementation of the path, run as the checker — TEST INFRASTRUCTURE.

The reference's hot path is ``DDP(model, device_ids=[gpu_id])`` + ``loss.backward()``
(ref:multigpu.py:36,42): stock ``torch.nn.parallel.DistributedDataParallel``.  This module runs exactly
that — on gloo/CPU anywhere (the reference's CPU-runnable configuration, BASELINE configs[0]) or on
nccl/GPU on the B200 box (the parity target north_star names) — for a seeded model and seeded per-rank
batches, and returns what a parity test needs: the gradients DDP leaves in ``param.grad`` at every step,
and the per-rank LOCAL gradients of the same batches (from an un-wrapped replica) so the closed-form
oracle (allreduce_oracle.ddp_average) can be pinned against it.

It is also the ``--impl reference`` arm of bench.py (timed with perf_counter on the host cores) and the
``cpu_baseline`` of the product arm.
"""
from __future__ import annotations

import copy
import os
import socket
import time
from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, spec: Dict, ret) -> None:
    backend = spec["backend"]
    # when the caller itself runs under torchrun, its TORCHELASTIC_USE_AGENT_STORE=True would make c10d look
    # for the elastic agent's store at OUR port instead of starting a store daemon on rank 0: this world is
    # independent of the launcher's, so the launcher's rendezvous variables are dropped
    for var in ("TORCHELASTIC_USE_AGENT_STORE", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_RUN_ID"):
        os.environ.pop(var, None)
    threads = spec.get("threads_per_rank")
    if threads:
        torch.set_num_threads(int(threads))
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        device = torch.device("cuda", rank) if backend == "nccl" else torch.device("cpu")
        if device.type == "cuda":
            torch.cuda.set_device(device)
        import b200ddp  # noqa: F401  (model/batch definitions only; the product path is NOT used here)
        from b200ddp import workloads

        wl = workloads.get(spec["workload"])
        build, make_batch, loss_fn = wl.build, wl.batch, wl.loss
        torch.manual_seed(spec["seed"])  # identical initial weights on every rank
        model = build().to(device)
        shadow = copy.deepcopy(model) if spec.get("record") else None
        ddp = torch.nn.parallel.DistributedDataParallel(
            model, device_ids=[rank] if device.type == "cuda" else None,
            **({"bucket_cap_mb": spec["bucket_cap_mb"]} if spec.get("bucket_cap_mb") else {}))
        opt = torch.optim.SGD(model.parameters(), lr=spec.get("lr", 1e-3))
        gen = torch.Generator().manual_seed(spec["seed"] + 1000 + rank)
        bsz = spec["batch_size"]
        rec: Dict[str, List] = {"ddp_grads": [], "local_grads": [], "weights": [], "loss": []}
        steps, warmup = spec["steps"], spec.get("warmup", 0)
        fixed = make_batch(bsz, device, gen) if spec.get("fixed_batch") else None
        # timing runs draw from a pre-generated pool so that batch synthesis stays outside the step,
        # as in the product arm (recording runs keep one fresh batch per step)
        pool = None
        if not spec.get("record") and fixed is None:
            pool = [make_batch(bsz, device, gen) for _ in range(min(warmup + steps, 64))]
        t0 = None
        for step in range(warmup + steps):
            if step == warmup:
                if device.type == "cuda":
                    torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
            if pool is not None:
                x, y = pool[step % len(pool)]
            else:
                x, y = fixed if fixed is not None else make_batch(bsz, device, gen)
            if shadow is not None:
                shadow.load_state_dict(model.state_dict())
                shadow.zero_grad()
                loss_fn(shadow(x), y).backward()
                rec["local_grads"].append([p.grad.detach().cpu().clone() for p in shadow.parameters()])
            opt.zero_grad()
            loss = loss_fn(ddp(x), y)
            loss.backward()
            if shadow is not None:
                rec["ddp_grads"].append([p.grad.detach().cpu().clone() for p in model.parameters()])
                rec["loss"].append(float(loss.detach()))
            opt.step()
            if spec.get("read_loss"):
                loss.item()
        if device.type == "cuda":
            torch.cuda.synchronize()
        dist.barrier()
        elapsed = time.perf_counter() - t0
        if shadow is not None:
            rec["weights"] = [p.detach().cpu().clone() for p in model.parameters()]
            try:
                log = ddp._get_ddp_logging_data()
                rec["bucket_sizes"] = log.get("bucket_sizes")
                rec["rebuilt_bucket_sizes"] = log.get("rebuilt_bucket_sizes")
            except Exception:  # logging data is a private API
                pass
        ret[rank] = {"elapsed_s": elapsed, "record": rec if shadow is not None else None}
    finally:
        dist.destroy_process_group()


def run_stock_ddp(world: int, workload: str, *, backend: str = "gloo", steps: int = 3, warmup: int = 0, batch_size: int = 32, seed: int = 1234, record: bool = True,
                  bucket_cap_mb: Optional[float] = None, threads_per_rank: Optional[int] = None,
                  fixed_batch: bool = False, read_loss: bool = False, lr: float = 1e-3) -> Dict[int, Dict]:
    """Spawn ``world`` processes running stock torch-DDP on ``workload`` (a name in
    distributed-pytorch_b200/workloads.py); returns {rank: {"elapsed_s", "record"}}."""
    spec = dict(backend=backend, workload=workload, steps=steps, warmup=warmup,
                batch_size=batch_size, seed=seed, record=record, bucket_cap_mb=bucket_cap_mb,
                threads_per_rank=threads_per_rank, fixed_batch=fixed_batch, read_loss=read_loss, lr=lr)
    port = free_port()
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, spec, ret), nprocs=world, join=True)
        return {r: dict(ret[r]) for r in range(world)}
