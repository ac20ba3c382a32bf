"""BASELINE configs[2], model level: the `slices` model (64 x Linear(4096, 64): 64 MiB of gradients in ~1 MiB tensors)
under DDP with bucket_cap_mb in {1, 2, 4, ..., 256} — samples/s and whole-bucket bus GB/s of our path, stock
torch-DDP/NCCL beside it, one JSON line per cap (rank 0).  Run under torchrun:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/cap_sweep.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import b200ddp  # noqa: E402,F401
import bench  # noqa: E402
from b200ddp import multigpu_torchrun, workloads  # noqa: E402

caps = [float(c) for c in os.environ.get("CAPS", "1,2,4,8,16,32,64,128,256").split(",")]
os.environ.setdefault("B200DDP_TIMEOUT_MS", "60000")
multigpu_torchrun.ddp_setup()
rank, world, local = dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
h = bench.Harness(torch, dist, dev, world, rank)
wl = workloads.get("slices")
bsz = wl.per_gpu_batch
host, devb, _ = bench.make_batches(torch, wl, "slices", bsz, local, rank)
steps, warmup, blocks = 40, 5, 3
for cap in caps:
    trainer = bench.build_trainer(torch, wl, "slices", bsz, local, rank, False, cap, False, "auto")
    value, e2e = bench.time_training(h, trainer, host, devb, None, steps, warmup, blocks, bsz, resident_ok=False)
    roof = bench.bucket_roofline(torch, h, trainer.model, 6551.0, "MEASURED_PEAKS.json", 900.0, "nominal NVLink 5")
    sizes = trainer.model.reducer.bucket_sizes()
    algos = [("fused" if b.fused else b.algo) for b in trainer.model.reducer.buckets]
    h.barrier_sync()
    trainer._graph = None
    trainer.model.shutdown()
    del trainer
    torch.cuda.empty_cache()
    stock = bench.stock_ddp_leg(torch, h, wl, "slices", bsz, local, rank, devb, steps, warmup, blocks, cap, False)
    if rank == 0:
        print(json.dumps({"bucket_cap_mb": cap, "n_gpus": world, "samples_s": value["value"], "ms_per_step": value["ms_per_step"],
                          "mode": value["mode"], "e2e_samples_s": e2e["value"], "buckets": len(sizes),
                          "largest_bucket_bytes": max(sizes), "algos": sorted(set(map(str, algos))),
                          "largest_bucket_ms": roof["avg_launch_ms"], "largest_bucket_bus_gbs": roof["achieved"],
                          "largest_bucket_frac": roof["frac"], "nccl_ddp_samples_s": stock["value"],
                          "nccl_ddp_ms_per_step": stock["ms_per_step"]}), flush=True)
multigpu_torchrun.destroy_process_group()
