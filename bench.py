#!/usr/bin/env python
"""bench.py — samples/sec of the reference's training step through the B200-native gradient-allreduce
path, plus the grad-allreduce GB/s-vs-bucket-size curve (BASELINE.json `metric`).

    python bench.py --gpus 1 --steps K --warmup W                      # product arm, one GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference stack on the host cores

A "step" is one `Trainer._run_batch` (ref:multigpu_torchrun.py:42-47: zero_grad, forward, loss, backward —
inside which the gradient buckets are flattened, all-reduced and unflattened by our kernels — and SGD
step) on one synthetic batch of 32 rows per GPU.  Default workload = BASELINE configs[1]: the toy
`Linear(20, 1)` on `MyTrainDataset`-shaped rows, 32 k rows (`--workload` selects wide / slices / resnet50 /
gpt2m).  The scripts' CrossEntropyLoss has identically zero gradients on this model, so the
measured step uses the multinode script's MSELoss (ref:multinode_torchrun.py:46): same kernels, same
bytes, non-vacuous parity.

One JSON line is printed by rank 0:
  value        whole-job samples/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          the same through the Trainer surface with HOST (pinned) batches: H2D of every batch and a
               D2H read of the loss inside the timed region
  roofline     the dominant kernel of the timed step (per-launch CUDA events on the launching stream)
  bucket_sweep kernel GB/s vs bucket size: flatten/unflatten against HBM at N=1, allreduce bus GB/s
               against NVLink at N>1 (the second half of BASELINE's metric)
  cpu_baseline stock torch-DDP on gloo/CPU (the reference stack) on a bounded sample, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD_DESC = {
    "toy": "configs[1]: multigpu_torchrun.py toy Linear(20,1), MyTrainDataset-shaped rows (32k), batch 32/GPU, "
           "SGD lr 1e-3, MSELoss (the scripts' CE has all-zero grads)",
    "wide": "configs[2]: Linear(4096,4096) on 4096-dim rows, batch 32/GPU (one 64 MiB gradient bucket)",
    "slices": "configs[2]: 64 x Linear(4096,64) heads, batch 32/GPU (64 MiB of gradients in ~1 MiB tensors)",
    "resnet50": "configs[3]: torchvision resnet50 random init, synthetic 3x224x224, batch 32/GPU, soft-label CE",
    "gpt2m": "configs[4]: GPT-2-medium-shaped transformer bf16, seq 1024, batch 8/GPU",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "nccl"],
                    help="b200: this repo; reference: the reference stack on the host cores (gloo/CPU); "
                         "nccl: stock torch-DDP over NCCL on the same GPUs (comparator, same loop)")
    ap.add_argument("--workload", default=os.environ.get("B200DDP_BENCH_WORKLOAD", "toy"), choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (auto: on for the launch-bound toy workload)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the bucket-size sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-max-mib", type=int, default=256)
    return ap.parse_args()


def default_steps(workload: str):
    return {"toy": (2000, 50), "wide": (300, 20), "slices": (200, 20), "resnet50": (30, 5), "gpt2m": (10, 3)}[workload]


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def mark(self):
        """Wall-clock bounds of the timed region; samples outside are dropped."""
        self.marks = getattr(self, "marks", []) + [time.time()]

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        import datetime

        marks = getattr(self, "marks", [])
        rows = out.splitlines()
        in_window = 0
        for use_window in (True, False):  # if no sample fell inside the (possibly very short) timed window, use
            if sm:                        # every sample of the run (taken under the same load) rather than none
                break
            for line in rows:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                if use_window and len(marks) >= 2:
                    try:
                        ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except ValueError:
                        ts = None
                    if ts is not None and not (marks[0] - 0.15 <= ts <= marks[-1] + 0.15):
                        continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            in_window = len(sm) if use_window else in_window
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "samples_in_timed_window": in_window}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: stock torch-DDP on gloo, host cores
# ------------------------------------------------------------------------------------------------------
def reference_run(workload: str, world: int, steps: int, warmup: int, read_loss: bool = True):
    """Times the reference stack's own implementation of the path (oracle/ref_ddp.py) on the host cores."""
    from oracle.ref_ddp import run_stock_ddp

    import b200ddp  # noqa: F401
    from b200ddp import workloads

    wl = workloads.get(workload)
    cores = os.cpu_count() or 1
    threads = max(1, cores // world)
    out = run_stock_ddp(world, workload, backend="gloo", steps=steps, warmup=warmup, batch_size=wl.per_gpu_batch,
                        record=False, threads_per_rank=threads, read_loss=read_loss)
    elapsed = max(out[r]["elapsed_s"] for r in out)
    sps = wl.per_gpu_batch * world * steps / elapsed
    return dict(value=sps, elapsed_s=elapsed, cores=min(cores, threads * world), threads_per_rank=threads,
                per_gpu_batch=wl.per_gpu_batch)


def reference_budget(workload: str, steps, warmup):
    """Bounded sample: about 10-30 s of CPU work per run."""
    cap = {"toy": (3000, 100), "wide": (40, 5), "slices": (40, 5), "resnet50": (3, 1), "gpt2m": (1, 1)}[workload]
    s = min(steps, cap[0]) if steps else cap[0]
    w = min(warmup, cap[1]) if warmup is not None else cap[1]
    return max(1, s), max(0, w)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = args.gpus
    steps, warmup = reference_budget(args.workload, args.steps, args.warmup)
    r = reference_run(args.workload, world, steps, warmup)
    line = {
        "impl": "reference", "metric": "samples/sec", "value": r["value"], "unit": "samples/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["elapsed_s"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.workload == "gpt2m" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "global_batch": r["per_gpu_batch"] * world,
                   "parallelism": f"dp{world} (stock torch-DDP, gloo, CPU)"},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{steps} steps after {warmup} warm-up: stock torch.nn.parallel.DistributedDataParallel "
                                   f"(the reference's dependency, unmodified) on gloo, {world} CPU process(es) x "
                                   f"{r['threads_per_rank']} threads, Trainer loop restated in oracle/ref_ddp.py"},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------------
def main_b200(args):
    import torch
    import torch.distributed as dist

    import b200ddp  # noqa: F401
    from b200ddp import _lib, multigpu_torchrun, workloads
    from b200ddp.comm import P2PComm, TensorTable

    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the b200ddp path has no CPU fallback"})
        return 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} must be launched with torch.distributed.run --nproc-per-node {args.gpus}")
    os.environ.setdefault("B200DDP_TIMEOUT_MS", "60000")  # a benchmark run should fail fast if a rank dies
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if "MASTER_PORT" not in os.environ:
        import socket

        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            os.environ["MASTER_PORT"] = str(sock.getsockname()[1])
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    steps, warmup = default_steps(args.workload)
    steps = args.steps if args.steps else steps
    warmup = max(3, args.warmup if args.warmup is not None else warmup)

    multigpu_torchrun.ddp_setup()  # init_process_group("nccl") + bind to LOCAL_RANK's GPU
    dev = torch.device("cuda", local)
    wl = workloads.get(args.workload)
    bsz = wl.per_gpu_batch

    # ---- objects, through the reference-facing surface ---------------------------------------------
    torch.manual_seed(1234)
    model = wl.build()
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-3)
    multigpu_torchrun.Trainer.loss_fn = staticmethod(wl.loss)
    host_batches = []
    if args.workload == "toy":
        # configs[1]: 32k MyTrainDataset rows, sharded by DistributedSampler exactly as prepare_dataloader
        # does (ref:multigpu.py:72-79), collated into pinned batches of 32
        from torch.utils.data.distributed import DistributedSampler

        from b200ddp.utils import DeviceShardLoader, MyTrainDataset

        ds = MyTrainDataset(32768, seed=1234)
        sampler = DistributedSampler(ds)
        idx = torch.tensor(list(sampler), dtype=torch.long)
        for i in range(0, len(idx) - bsz + 1, bsz):
            sel = idx[i:i + bsz]
            host_batches.append((ds.x[sel].pin_memory(), ds.y[sel].pin_memory()))
        # the HBM-resident leg reads the same shard through the device-resident loader
        shard_loader = DeviceShardLoader(ds, bsz, torch.device("cuda", local), sampler=sampler)
    else:
        gen = torch.Generator().manual_seed(4321 + rank)
        for _ in range(8):
            x, y = wl.batch(bsz, torch.device("cpu"), gen)
            host_batches.append((x.pin_memory(), y.pin_memory()))
    n_host_batches = len(host_batches)
    stock = args.impl == "nccl"
    if stock:  # comparator: the reference's own binding of the name DDP (ref:multigpu_torchrun.py:7)
        multigpu_torchrun.DDP = torch.nn.parallel.DistributedDataParallel
    trainer = multigpu_torchrun.Trainer(model, host_batches, optimizer, save_every=10**9,
                                        snapshot_path="/nonexistent/snapshot.pt")
    ddp = trainer.model

    class _NoComm:  # stock DDP has no b200ddp communicator
        launch_count = 0
        mem_mode = "nccl"
        has_multicast = None

    comm = _NoComm() if stock else ddp.engine.comm
    if args.workload == "toy":
        dev_batches = [(x, y) for x, y in shard_loader if x.shape[0] == bsz][:256]
    else:
        dev_batches = [(x.to(dev), y.to(dev)) for x, y in host_batches[: min(len(host_batches), 256)]]
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])

    def barrier_sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ---------------------------------------------------------------
    use_graph = (args.graph == "on" or (args.graph == "auto" and args.workload == "toy")) and not stock
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_steps(n):
        """EXACTLY n steps bracketed by barrier + synchronize; device time, max over ranks."""
        barrier_sync()
        e0.record()
        for i in range(n):
            trainer._run_batch(*dev_batches[i % len(dev_batches)])
        e1.record()
        barrier_sync()
        return max_over_ranks(e0.elapsed_time(e1))

    clocks = ClockSampler(local).start() if rank == 0 else None
    for i in range(warmup):
        trainer._run_batch(*dev_batches[i % len(dev_batches)])
    eager = None
    if use_graph:
        # eager number first (for the record), then capture: zero_grad/forward/loss/backward with the bucket
        # kernels/SGD step become ONE graph launch; every replay does the same work on a new batch
        eager_ms = timed_steps(steps)
        eager = {"value": bsz * world * steps / (eager_ms / 1e3), "ms_per_step": eager_ms / steps,
                 "note": "same step launched eagerly (no CUDA graph)"}
        trainer.enable_cuda_graph(*dev_batches[0])
        for i in range(warmup):
            trainer._run_batch(*dev_batches[i % len(dev_batches)])
    barrier_sync()
    if clocks:
        clocks.mark()
    n0 = comm.launch_count
    ms_total = timed_steps(steps)
    if clocks:
        clocks.mark()
    launches = (trainer.graph_kernel_launches * steps) if use_graph else (comm.launch_count - n0)
    clock_info = clocks.stop() if clocks else None
    value = bsz * world * steps / (ms_total / 1e3)

    # ---- e2e: host batches through the Trainer surface, loss read back every step ----------------------
    e2e_steps = steps
    for i in range(min(warmup, 10)):
        x, y = host_batches[i % n_host_batches]
        trainer._run_batch(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)).item()
    barrier_sync()
    e0.record()
    for i in range(e2e_steps):
        x, y = host_batches[i % n_host_batches]
        if use_graph:  # as Trainer._run_epoch does: pinned host batch -> the graph's input buffers (H2D) -> replay
            loss = trainer._run_batch(x, y)
        else:
            loss = trainer._run_batch(x.to(trainer.gpu_id, non_blocking=True), y.to(trainer.gpu_id, non_blocking=True))
        loss.item()
    e1.record()
    barrier_sync()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": bsz * world * e2e_steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps}

    if stock:
        if rank == 0:
            emit({"impl": "nccl", "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
                  "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
                  "scaling": "weak", "data": "synthetic", "e2e": e2e, "clocks": clock_info,
                  "config": {"workload": WORKLOAD_DESC[args.workload], "global_batch": bsz * world,
                             "parallelism": f"dp{world} (stock torch-DDP + NCCL, comparator)"}})
        dist.destroy_process_group()
        return 0
    # ---- roofline of the step's dominant kernel: per-launch events on the launching stream -------------
    roof_steps = min(steps, 200)
    trainer._graph = None  # the per-launch events need eager launches of the same kernels
    engine = ddp.engine
    pairs = []
    orig_reduce = engine.reduce_bucket

    def timed_reduce(bucket, grads):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        comm.stream_wait(engine._stream_ptr, torch.cuda.current_stream(dev).cuda_stream)
        a.record(engine.stream)
        orig_reduce(bucket, grads)
        b.record(engine.stream)
        pairs.append((bucket.index, a, b))

    engine.reduce_bucket = timed_reduce
    for i in range(roof_steps):
        trainer._run_batch(*dev_batches[i % len(dev_batches)])
    barrier_sync()
    engine.reduce_bucket = orig_reduce
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    per_bucket = {}
    for bi, a, b in pairs:
        per_bucket.setdefault(bi, []).append(a.elapsed_time(b))
    buckets = {b.index: b for b in ddp.reducer.buckets}
    dom = max(per_bucket, key=lambda k: statistics.mean(per_bucket[k]))
    dom_ms = statistics.mean(per_bucket[dom])
    S = buckets[dom].nbytes
    if world == 1:
        # flatten (read S, write S) + unflatten (read S, write S); the fused kernel reads S, writes S twice
        alg_bytes = (3 if buckets[dom].fused else 4) * S
        bound, peak, unit = "hbm", hbm_peak, "GB/s"
        what = "fused_bucket_kernel" if buckets[dom].fused else "multi_tensor_copy_kernel x2 (flatten+unflatten)"
    else:
        alg_bytes = S * 2 * (world - 1) / world  # NVLink bytes per GPU and direction (bus-bandwidth convention)
        bound, peak, unit = "nvlink", 900.0, "GB/s"
        peak_src = "nominal NVLink 5: 900 GB/s per direction per GPU (measured peer copy on this pool: 770 GB/s)"
        what = "fused_bucket_kernel" if buckets[dom].fused else "flatten + allreduce + unflatten of the largest bucket"
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": None, "kernel": what, "bucket_bytes": S, "avg_launch_ms": dom_ms,
                "launches_timed": len(per_bucket[dom]), "peak_source": peak_src,
                "note": "dominant bucket of the timed step; bucket_sweep holds the size curve"}

    # ---- bucket-size sweep ------------------------------------------------------------------------------
    sweep = []
    if not args.no_sweep:
        sweep = bucket_sweep(torch, dist, comm_factory=lambda nbytes: P2PComm.from_process_group(nbytes, local),
                             world=world, rank=rank, dev=dev, hbm_peak=hbm_peak, max_mib=args.sweep_max_mib,
                             lib=_lib, TensorTable=TensorTable)
    # ---- cpu baseline (N=1, rank 0) ----------------------------------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        s, w = reference_budget(args.workload, None, None)
        r = reference_run(args.workload, 1, s, w)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
               "sample": f"{s} steps (batch {bsz}) after {w} warm-up of stock torch-DDP on gloo/CPU, 1 process x "
                         f"{r['threads_per_rank']} threads, via oracle/ref_ddp.py"}

    line = {
        "impl": "b200", "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if wl.dtype == torch.bfloat16 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "global_batch": bsz * world,
                   "parallelism": f"dp{world}", "bucket_bytes": ddp.reducer.bucket_sizes(),
                   "bucket_algos": [("fused" if b.fused else b.algo) for b in ddp.reducer.buckets],
                   "mem_mode": comm.mem_mode, "multicast": comm.has_multicast, "cuda_graph": use_graph,
                   "l2": "training steps are not L2-flushed (each step consumes a different batch; weights are "
                         "the step's own state); bucket_sweep rotates buffers totalling > 126 MB L2"},
        "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "bucket_sweep": sweep, "cpu_baseline": cpu, "eager": eager,
    }
    # `roofline` reports the path's dominant kernel where it is bandwidth-bound: the largest bucket of the
    # sweep measured in this same run (flatten at N=1 against HBM, the AUTO allreduce at N>1 against NVLink).
    # The step's own dominant bucket stays in `roofline_step` (for the toy workload it is 84 bytes: pure
    # latency, its fraction of any bandwidth roofline is ~0 by construction).
    line["roofline_step"] = roofline
    if sweep:
        if world == 1:
            cand = [p for p in sweep if p["kernel"] == "flatten_scale" and p["bytes"] == (64 << 20)] or \
                   [p for p in sweep if p["kernel"] == "flatten_scale"]
        else:
            ours = [p for p in sweep if p.get("frac") is not None]
            top = max(p["bytes"] for p in ours)
            cand = [max((p for p in ours if p["bytes"] == top), key=lambda p: p["achieved"])]
        best = cand[-1]
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tr.get(f"{best['kernel']}@{best['bytes']}", {}).get("dram_bytes_per_launch")
        except OSError:
            pass
        line["roofline"] = {"bound": best["bound"], "achieved": best["achieved"], "peak": best["peak"],
                            "unit": "GB/s", "frac": best["frac"], "traffic": traffic, "kernel": best["kernel"],
                            "bucket_bytes": best["bytes"], "avg_launch_ms": best["ms"],
                            "algorithmic_bytes_per_launch": (2 * best["bytes"] if world == 1 else
                                                             best["bytes"] * 2 * (world - 1) / world),
                            "peak_source": (peak_src if world == 1 else
                                            "nominal NVLink 5: 900 GB/s per direction per GPU (measured peer copy on "
                                            "this pool: 770 GB/s)"),
                            "where": "bucket_sweep of this run (CUDA-graph replay timed with CUDA events on the "
                                     "replaying stream); the timed step's own bucket is in roofline_step"}
    if rank == 0:
        emit(line)
    multigpu_torchrun.destroy_process_group()
    return 0


def bucket_sweep(torch, dist, comm_factory, world, rank, dev, hbm_peak, max_mib, lib, TensorTable):
    """Kernel time vs bucket size.  `iters` launches (rotating through > L2 of distinct buffers) are captured
    into ONE CUDA graph and the replay is timed with CUDA events on the replaying stream: per-launch time =
    replay time / iters.  This keeps the host (ctypes launch + event records, and the skew between the
    ranks' Python loops) out of a microsecond-scale measurement; the launches are graph-capturable because
    the kernels keep their barrier generation in device memory.
    N=1: flatten and unflatten (HBM roofline, algorithmic bytes 2 S each).  N>1: each allreduce algorithm
    (bus GB/s = S * 2 (ws-1)/ws / t against 900 GB/s per direction) and NCCL on the same box."""
    sizes = [m << 20 for m in (1, 4, 16, 64, 256) if m <= max_mib]
    if world > 1:
        sizes = [256 << 10] + sizes
    biggest = max(sizes)
    slots = max(2, (288 << 20) // biggest + 1)  # rotate through >= 288 MiB of distinct arena
    arena = biggest * slots + (8 << 20)
    comm = comm_factory(arena)
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    out = []

    def time_graph(fn, iters, replays=5):
        """median over `replays` of (graph replay time / iters), max over ranks, in ms"""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(iters):
                fn(i)
        ts = []
        with torch.cuda.stream(stream):
            for rep in range(replays + 2):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                g.replay()
                b.record(stream)
                stream.synchronize()
                if rep >= 2:
                    ts.append(a.elapsed_time(b) / iters)
        ms = statistics.median(ts)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        del g
        return ms

    def iters_for(S):
        return 40 if S <= (4 << 20) else (20 if S <= (64 << 20) else 8)

    if world == 1:
        for S in sizes:
            n_t = 8
            per = S // 4 // n_t
            reps = max(2, (288 << 20) // S + 1)
            srcs = [[torch.rand(per, device=dev) for _ in range(n_t)] for _ in range(reps)]
            offs = [i * per for i in range(n_t)]
            tables = [TensorTable([t.data_ptr() for t in ts], offs, [per] * n_t) for ts in srcs]
            # both the gradient side and the bucket side rotate through > L2 of distinct memory
            for name, fn in (("flatten_scale", lambda i: comm.flatten_scale(tables[i % reps], lib.F32,
                                                                             (i % (reps - 1)) * S, 0.5, sp)),
                             ("unflatten", lambda i: comm.unflatten(tables[i % reps], lib.F32,
                                                                     (i % (reps - 1)) * S, sp))):
                ms = time_graph(fn, max(iters_for(S), reps))  # one replay streams > L2 of distinct buffers
                gbs = 2 * S / (ms * 1e-3) / 1e9
                out.append({"kernel": name, "bytes": S, "ms": ms, "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                            "frac": gbs / hbm_peak, "unit": "GB/s"})
            del srcs, tables
    else:
        algos = [("oneshot", 1), ("twoshot", 2), ("twoshot_tma", 3)] + ([("nvls", 4)] if comm.has_multicast else [])
        comm.arena().view(torch.float32).uniform_(-1e-3, 1e-3)
        torch.cuda.synchronize()
        dist.barrier()
        for S in sizes:
            reps = max(2, (288 << 20) // S + 1)
            for name, algo in algos:
                if name == "oneshot" and S > (16 << 20):
                    continue

                def fn(i, S=S, algo=algo, name=name, reps=reps):
                    off = (i % (reps - 1)) * S if name != "oneshot" else 0
                    comm.allreduce(off, off if name != "oneshot" else (reps - 1) * S, S, lib.F32, algo, 0, sp)

                ms = time_graph(fn, max(iters_for(S), reps - 1))
                bus = S * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
                out.append({"kernel": f"allreduce_{name}", "bytes": S, "ms": ms, "bound": "nvlink", "achieved": bus,
                            "peak": 900.0, "frac": bus / 900.0, "unit": "GB/s (bus)", "algbw": S / (ms * 1e-3) / 1e9})
                comm.poll_error()
        # the comparator on the same box: NCCL all_reduce (what stock DDP calls), timed the same way
        for S in sizes:
            reps = max(2, min(8, (288 << 20) // S + 1))
            bufs = [torch.empty(S // 4, device=dev).uniform_(-1e-3, 1e-3) for _ in range(reps)]
            for b in bufs[:2]:
                dist.all_reduce(b)
            method = "cuda graph"
            try:
                with torch.cuda.stream(stream):
                    dist.all_reduce(bufs[0])  # make sure NCCL has used this stream before capture
                ms = time_graph(lambda i: dist.all_reduce(bufs[i % reps]), iters_for(S))
            except Exception as exc:  # capture of NCCL can be refused by some builds: time eagerly instead
                method = f"eager loop ({type(exc).__name__})"
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = iters_for(S)
                a.record()
                for i in range(n):
                    dist.all_reduce(bufs[i % reps])
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / n
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            out.append({"kernel": "nccl_all_reduce (comparator, not ours)", "bytes": S, "ms": ms, "bound": "nvlink",
                        "achieved": S * 2 * (world - 1) / world / (ms * 1e-3) / 1e9, "peak": 900.0, "frac": None,
                        "unit": "GB/s (bus)", "timed": method})
            del bufs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    comm.destroy()
    return out


def main():
    args = parse_args()
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner) and child
    # processes write to fd 1 too, so fd 1 is pointed at stderr and the line goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
