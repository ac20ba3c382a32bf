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

The Trainer is used the way the scripts use it: its default (B200DDP_GRAPH=auto) replays the step from a CUDA
graph after three eager steps; `--graph off` forces every step eager.

One JSON line is printed by rank 0:
  value         whole-job samples/s, inputs resident in HBM (toy: the rank's sampler shard, gathered by a
                device-side cursor inside the replayed step), CUDA-event timed, max over ranks; the timed region is
                `timed_blocks` blocks of EXACTLY `steps` steps, each bracketed by barrier + synchronize; the
                reported time is the median block (a 20-step block of the toy step lasts ~0.5 ms: one sample
                cannot resolve a 5 % change)
  e2e           the same through `Trainer._run_batch` with HOST (pinned) batches: H2D of every batch and a D2H read
                of every step's loss inside the timed region (the replayed step begins with its own H2D copy and
                ends with the loss D2H; the loss of step i is read while step i+1 runs)
  eager         the same step with B200DDP_GRAPH=0 (every kernel launched from Python/C++ hooks)
  nccl_ddp      stock torch-DDP over NCCL on the same GPUs, same loop, same batches (the same-hardware comparator)
  roofline      the dominant bandwidth-bound kernel at the bucket size DDP produces (N=1: flatten of a
                ResNet-50-shaped 25 MiB bucket against HBM; N>1: the AUTO allreduce at 25 MiB against NVLink)
  roofline_step the timed step's own largest bucket: all its kernels (flatten+allreduce+unflatten, or the one fused
                launch) replayed back to back from a CUDA graph, device time per bucket
  bucket_sweep  kernel GB/s vs bucket size (the second half of BASELINE's metric), NCCL beside it at N>1
  parity        N>1: every algorithm AUTO can pick, ragged sizes, f32 and bf16, against NCCL and a rank-ordered sum
  nvlink        N>1: measured peer-read bandwidth and NVML NVLink byte counters around the allreduce kernels
  extra         toy runs only: the same step on ResNet-50 (configs[3]: realistic ~100 MB gradient volume)
  cpu_baseline  stock torch-DDP on gloo/CPU (the reference stack) on a bounded sample, N=1 only
"""
from __future__ import annotations

import argparse
import json
import math
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
L2_NOTE = ("training steps are not L2-flushed: every step consumes a different batch and the weights are the step's "
           "own state; the kernel sweeps rotate buffers totalling > 126 MB L2")
NVLINK_NOMINAL = 900.0  # GB/s per direction per GPU
NVLINK_POOL_MEASURED = 770.0  # B200_PROFILING.md: peer copy measured on this pool


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
                    help="auto: the Trainer's default (replay after three eager steps); off: every step eager")
    ap.add_argument("--blocks", type=int, default=None, help="timed blocks of `steps` steps (default: by workload)")
    ap.add_argument("--bucket-cap-mb", type=float, default=None, help="DDP bucket_cap_mb (the slices sweep)")
    ap.add_argument("--views", action="store_true", help="gradient_as_bucket_view=True (both our DDP and the comparator)")
    ap.add_argument("--fuse-sgd", action="store_true",
                    help="opt-in optimizer epilogue (B200DDP_FUSE_SGD=1): the bucket kernels also apply the SGD step")
    ap.add_argument("--no-sweep", action="store_true", help="skip the bucket-size sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the ResNet-50 leg of toy runs")
    ap.add_argument("--no-nccl", action="store_true", help="skip the stock-DDP/NCCL comparator leg")
    ap.add_argument("--sweep-max-mib", type=int, default=256)
    return ap.parse_args()


def default_steps(workload: str):
    return {"toy": (2000, 50), "wide": (300, 20), "slices": (200, 20), "resnet50": (30, 5), "gpt2m": (10, 3)}[workload]


def blocks_for(workload: str, steps: int, asked):
    """Enough blocks of `steps` steps to cover the workload's default step count, at most 30."""
    if asked:
        return max(1, asked)
    total = default_steps(workload)[0]
    return max(1, min(30, math.ceil(total / max(1, steps))))


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
        self.marks = []

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
        self.marks.append(time.time())

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

        marks = self.marks
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
def bench_config(workload: str, world: int, per_gpu_batch: int) -> dict:
    """The SAME dict in the product, nccl and reference arms (implementation details live in `impl_config`)."""
    return {"workload": WORKLOAD_DESC[workload], "global_batch": per_gpu_batch * world, "parallelism": f"dp{world}",
            "l2": L2_NOTE}


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
        "data": "synthetic", "config": bench_config(args.workload, world, r["per_gpu_batch"]),
        "impl_config": {"stack": "stock torch.nn.parallel.DistributedDataParallel, gloo, CPU",
                        "processes": world, "threads_per_process": r["threads_per_rank"]},
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
class Harness:
    """Process-group, timing and reduction helpers shared by the legs."""

    def __init__(self, torch, dist, dev, world, rank):
        self.torch, self.dist, self.dev, self.world, self.rank = torch, dist, dev, world, rank
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def barrier_sync(self):
        self.torch.cuda.synchronize()
        self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_blocks(self, step_fn, steps: int, blocks: int, drain=None):
        """`blocks` blocks of EXACTLY `steps` calls of step_fn(i); each block is bracketed by barrier +
        synchronize on both sides and timed with CUDA events on the current stream; per block the MAX over
        ranks is taken; returns (median block ms, all block ms).  `drain()` runs inside the timed region after
        the last step of a block (the pipelined e2e leg reads its last loss there)."""
        out = []
        i = 0
        for _ in range(blocks):
            self.barrier_sync()
            self.e0.record()
            for _ in range(steps):
                step_fn(i)
                i += 1
            if drain is not None:
                drain()
            self.e1.record()
            self.barrier_sync()
            out.append(self.max_over_ranks(self.e0.elapsed_time(self.e1)))
        return statistics.median(out), out


def block_stats(ms_list, steps):
    return {"timed_blocks": len(ms_list), "block_ms_min": min(ms_list), "block_ms_median": statistics.median(ms_list),
            "block_ms_max": max(ms_list), "steps_per_block": steps}


def build_trainer(torch, wl, workload, bsz, local, rank, stock: bool, cap_mb, views: bool, graph_mode: str):
    """Objects through the reference-facing surface (multigpu_torchrun.Trainer)."""
    from b200ddp import multigpu_torchrun

    torch.manual_seed(1234)
    model = wl.build()
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-3)
    kw = {}
    if cap_mb is not None:
        kw["bucket_cap_mb"] = cap_mb
    if views:
        kw["gradient_as_bucket_view"] = True
    ours = multigpu_torchrun.DDP
    base = torch.nn.parallel.DistributedDataParallel if stock else ours
    # the scripts call DDP(model, device_ids=[gpu_id]); bench-only keyword arguments ride on the same call
    multigpu_torchrun.DDP = (lambda m, device_ids=None: base(m, device_ids=device_ids, **kw)) if kw else base
    multigpu_torchrun.Trainer.loss_fn = staticmethod(wl.loss)
    try:
        trainer = multigpu_torchrun.Trainer(model, [], optimizer, save_every=10**9,
                                            snapshot_path="/nonexistent/snapshot.pt")
    finally:
        multigpu_torchrun.DDP = ours
    st = trainer._state()
    st["mode"] = "off" if (stock or graph_mode == "off") else ("on" if graph_mode == "on" else "auto")
    return trainer


def make_batches(torch, wl, workload, bsz, local, rank):
    """(host pinned batches, device batches, shard loader or None)"""
    dev = torch.device("cuda", local)
    host, shard = [], None
    if workload == "toy":
        # configs[1]: 32k MyTrainDataset rows, sharded by DistributedSampler exactly as prepare_dataloader
        # does (ref:multigpu.py:72-79), collated into pinned batches of 32
        from torch.utils.data.distributed import DistributedSampler

        from b200ddp.utils import DeviceShardLoader, MyTrainDataset

        ds = MyTrainDataset(32768, seed=1234)
        sampler = DistributedSampler(ds)
        idx = torch.tensor(list(sampler), dtype=torch.long)
        for i in range(0, len(idx) - bsz + 1, bsz):
            sel = idx[i:i + bsz]
            host.append((ds.x[sel].pin_memory(), ds.y[sel].pin_memory()))
        # the HBM-resident leg reads the same shard through the device-resident loader
        shard = DeviceShardLoader(ds, bsz, dev, sampler=sampler)
        devb = [(x, y) for x, y in shard if x.shape[0] == bsz][:256]
    else:
        gen = torch.Generator().manual_seed(4321 + rank)
        for _ in range(8):
            x, y = wl.batch(bsz, torch.device("cpu"), gen)
            host.append((x.pin_memory(), y.pin_memory()))
        devb = [(x.to(dev), y.to(dev)) for x, y in host]
    return host, devb, shard


def time_training(h: Harness, trainer, host, devb, shard, steps, warmup, blocks, bsz, resident_ok: bool):
    """value (HBM-resident inputs) and e2e (host inputs, loss read every step) of one Trainer."""
    world = h.world
    st = trainer._state()
    nb = len(devb)
    use_resident = bool(resident_ok and shard is not None and st["mode"] != "off")
    for i in range(max(warmup, 4)):  # carries the eager steps + capture of the Trainer's auto mode
        trainer._run_batch(*devb[i % nb])
    if use_resident:
        use_resident = trainer.enable_resident_shard(shard)
    if use_resident:
        step = lambda i: trainer._run_resident_step()  # noqa: E731
    else:
        step = lambda i: trainer._run_batch(*devb[i % nb])  # noqa: E731
    for i in range(warmup):
        step(i)
    ms, all_ms = h.timed_blocks(step, steps, blocks)
    per_replay = (st["resident"]["launches_per_replay"] if use_resident else trainer.graph_kernel_launches)
    value = {"value": bsz * world * steps / (ms / 1e3), "ms_per_step": ms / steps, **block_stats(all_ms, steps),
             "launches_per_replay": per_replay,
             "mode": ("cuda-graph replay, batch gathered from the HBM-resident shard inside the replay" if use_resident
                      else "cuda-graph replay (device batch copied into the replay's input buffers)"
                      if st["slots"].get("dev") is not None else "eager")}
    # ---- e2e ----
    nh = len(host)
    pending = [None]

    def e2e_step(i):
        x, y = host[i % nh]
        out = trainer._run_batch(x, y)
        if pending[0] is not None:
            pending[0].item()  # the PREVIOUS step's loss: its D2H copy was part of that step's replay
        pending[0] = out

    def drain():
        if pending[0] is not None:
            pending[0].item()
            pending[0] = None

    for i in range(max(min(warmup, 10), 6)):
        e2e_step(i)
    drain()
    ems, eall = h.timed_blocks(e2e_step, steps, blocks, drain=drain)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    e2e = {"value": bsz * world * steps / (ems / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": ems / steps, "steps": steps, **block_stats(eall, steps),
           "mode": ("two alternating cuda-graph replays, each = H2D(pinned staging) + step + D2H(loss); loss of step "
                    "i read after step i+1 is queued" if st["slots"].get("host") is not None else
                    "eager: x.to(gpu) + step + loss.item() per step")}
    return value, e2e


def bucket_roofline(torch, h: Harness, ddp, hbm_peak, peak_src, nvl_peak, nvl_src):
    """Device time of the step's LARGEST bucket, all its kernels back to back: `iters` x engine.reduce_bucket
    captured into one CUDA graph, replay timed with CUDA events / iters (collective: every rank replays)."""
    engine, red = ddp.engine, ddp.reducer
    b = max(red.buckets, key=lambda k: k.nbytes)
    grads = [red.params[i].grad if red.params[i].grad is not None else torch.zeros_like(red.params[i])
             for i in b.param_indices]
    iters = 20 if b.nbytes <= (32 << 20) else 6
    torch.cuda.synchronize()
    h.dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            engine.reduce_bucket(b, grads)
        engine.finalize()
    ts = []
    for rep in range(6):
        h.barrier_sync()
        h.e0.record()
        g.replay()
        h.e1.record()
        torch.cuda.synchronize()
        if rep >= 1:
            ts.append(h.e0.elapsed_time(h.e1) / iters)
    ms = h.max_over_ranks(statistics.median(ts))
    S = b.nbytes
    world = h.world
    if b.fused:
        kernels = "fused_ll_kernel" if (world > 1 and S <= (32 << 10)) else "fused_bucket_kernel"
        n_k = 1
    else:
        ks = ["multi_tensor_copy_kernel<flatten>"] + ([f"allreduce algo {b.algo}"] if world > 1 else []) + \
             ([] if engine.grad_views else ["multi_tensor_copy_kernel<unflatten>"])
        kernels, n_k = " + ".join(ks), len(ks)
    if world == 1:
        alg = (3 if b.fused else (2 if engine.grad_views else 4)) * S
        bound, peak, src = "hbm", hbm_peak, peak_src
        note = "algorithmic bytes: read S + write S per copy pass (fused: read S, write 2 S)"
    else:
        alg = S * 2 * (world - 1) / world
        bound, peak, src = "nvlink", nvl_peak, nvl_src
        note = "bus-bandwidth convention S*2(ws-1)/ws over the WHOLE bucket (flatten and unflatten included)"
    ach = alg / (ms * 1e-3) / 1e9
    del g
    if world == 1 and 3 * S < (100 << 20):
        note += ("; gradients and bucket stay L2-resident across the back-to-back replays (126 MB L2), as they largely "
                 "are in the real step (flatten reads what the backward just wrote, unflatten what flatten wrote): "
                 "a fraction above 1.0 of the HBM peak is L2 bandwidth, not DRAM")
    return {"bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "kernel": kernels, "kernels_per_bucket": n_k, "bucket_bytes": S, "tensors": len(b.param_indices),
            "avg_launch_ms": ms, "launches_timed": iters * 5, "peak_source": src, "note": note,
            "how": "CUDA-graph replay of back-to-back engine.reduce_bucket calls, CUDA events, max over ranks"}


def resnet_table(torch, TensorTable, dev, reps):
    """The largest gradient bucket torch's bucketing gives ResNet-50 (about 30 MiB: two 9 MiB convolutions, 1x1
    convolutions and dozens of 256 B - 8 KiB BatchNorm vectors), as flatten tables over `reps` sets of source tensors."""
    from torchvision.models import resnet50

    from b200ddp.bucketing import compute_bucket_assignment_by_size, layout_bucket

    with torch.device("meta"):
        m = resnet50()
    ps = [p for p in m.parameters()]
    sizes = [p.numel() * 4 for p in ps]
    idx, _ = compute_bucket_assignment_by_size(sizes, [0] * len(ps), [1 << 20, 25 << 20])
    big = max(idx, key=lambda b: sum(sizes[i] for i in b))
    numels = [ps[i].numel() for i in big]
    spec = layout_bucket(0, big, numels, 0, 4)
    tables = []
    keep = []
    for _ in range(reps):
        ts = [torch.rand(n, device=dev) for n in numels]
        keep.append(ts)
        tables.append(TensorTable([t.data_ptr() for t in ts], spec.offsets_el, spec.lengths_el))
    return spec.nbytes, len(numels), tables, keep


def bucket_sweep(torch, dist, comm_factory, world, rank, dev, hbm_peak, nvl_peak, max_mib, lib, TensorTable):
    """Kernel time vs bucket size.  `iters` launches (rotating through > L2 of distinct buffers) are captured
    into ONE CUDA graph and the replay is timed with CUDA events on the replaying stream: per-launch time =
    replay time / iters.  This keeps the host (ctypes launch + event records, and the skew between the
    ranks' Python loops) out of a microsecond-scale measurement; the launches are graph-capturable because
    the kernels keep their barrier generation in device memory.
    N=1: flatten and unflatten (HBM roofline, algorithmic bytes 2 S each), clean 8-tensor buckets and a
    ResNet-50-shaped one.  N>1: each allreduce algorithm (bus GB/s = S * 2 (ws-1)/ws / t against NVLink) and NCCL
    on the same box."""
    sizes = [m << 20 for m in (1, 4, 16, 25, 64, 256) if m <= max_mib]
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
        def copy_points(S, tables, reps, label, extra):
            # both the gradient side and the bucket side rotate through > L2 of distinct memory
            for name, fn in (("flatten_scale", lambda i: comm.flatten_scale(tables[i % reps], lib.F32,
                                                                             (i % (reps - 1)) * S_slot[0], 0.5, sp)),
                             ("unflatten", lambda i: comm.unflatten(tables[i % reps], lib.F32,
                                                                     (i % (reps - 1)) * S_slot[0], sp))):
                ms = time_graph(fn, max(iters_for(S), reps))  # one replay streams > L2 of distinct buffers
                gbs = 2 * S / (ms * 1e-3) / 1e9
                out.append({"kernel": name + label, "bytes": S, "ms": ms, "bound": "hbm", "achieved": gbs,
                            "peak": hbm_peak, "frac": gbs / hbm_peak, "unit": "GB/s", **extra})

        S_slot = [0]
        for S in sizes:
            n_t = 8
            per = S // 4 // n_t
            reps = max(2, (288 << 20) // S + 1)
            srcs = [[torch.rand(per, device=dev) for _ in range(n_t)] for _ in range(reps)]
            offs = [i * per for i in range(n_t)]
            tables = [TensorTable([t.data_ptr() for t in ts], offs, [per] * n_t) for ts in srcs]
            S_slot[0] = S
            copy_points(S, tables, reps, "", {"tensors": n_t})
            del srcs, tables
            # comparator, not ours: torch's own copy kernel (what MEASURED_PEAKS.json's hbm_gbs was measured with, at
            # 2 GiB) on ONE contiguous tensor of the same size in the same harness — what a plain copy reaches here
            a = [torch.rand(S // 4, device=dev) for _ in range(reps)]
            b = [torch.empty(S // 4, device=dev) for _ in range(reps)]
            ms = time_graph(lambda i: b[i % reps].copy_(a[i % reps]), max(iters_for(S), reps))
            gbs = 2 * S / (ms * 1e-3) / 1e9
            out.append({"kernel": "torch_copy (comparator, not ours)", "bytes": S, "ms": ms, "bound": "hbm",
                        "achieved": gbs, "peak": hbm_peak, "frac": gbs / hbm_peak, "unit": "GB/s", "tensors": 1})
            del a, b
        try:
            reps = 10
            S, n_t, tables, keep = resnet_table(torch, TensorTable, dev, reps)
            S_slot[0] = (S + 4095) // 4096 * 4096
            copy_points(S, tables, reps, "[resnet50 bucket]", {"tensors": n_t})
            del tables, keep
        except Exception as exc:  # torchvision missing: the clean points stand alone
            out.append({"kernel": "flatten_scale[resnet50 bucket]", "error": f"{type(exc).__name__}: {exc}"})
    else:
        algos = [("oneshot", 1), ("twoshot", 2), ("twoshot_tma", 3)] + ([("nvls", 4)] if comm.has_multicast else [])
        comm.arena().view(torch.float32).uniform_(-1e-3, 1e-3)
        torch.cuda.synchronize()
        dist.barrier()
        for S in sizes:
            reps = max(2, (288 << 20) // S + 1)
            auto = comm.pick_algo(S)
            for name, algo in algos:
                if name == "oneshot" and S > (16 << 20):
                    continue

                def fn(i, S=S, algo=algo, name=name, reps=reps):
                    off = (i % (reps - 1)) * S if name != "oneshot" else 0
                    comm.allreduce(off, off if name != "oneshot" else (reps - 1) * S, S, lib.F32, algo, 0, sp)

                ms = time_graph(fn, max(iters_for(S), reps - 1))
                bus = S * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
                out.append({"kernel": f"allreduce_{name}", "bytes": S, "ms": ms, "bound": "nvlink", "achieved": bus,
                            "peak": nvl_peak, "frac": bus / nvl_peak, "unit": "GB/s (bus)", "algbw": S / (ms * 1e-3) / 1e9,
                            "auto": algo == auto})
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
                        "achieved": S * 2 * (world - 1) / world / (ms * 1e-3) / 1e9, "peak": nvl_peak, "frac": None,
                        "unit": "GB/s (bus)", "timed": method})
            del bufs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    comm.destroy()
    return out


def parity_block(torch, dist, comm, world, rank, dev, lib, TensorTable):
    """Every algorithm AUTO can pick, at ragged sizes, f32 and bf16: our result against (a) `dist.all_reduce` (NCCL,
    what stock DDP calls) and (b) a rank-ordered fp32 sum of the all-gathered inputs (the order our peer-to-peer
    kernels use; NVLS lets the switch choose).  Returns {"algo/dtype": {"vs_nccl": max rel, "vs_ordered": max rel}}
    with rel = max|a-b| / max|b| over the buffer, max over sizes and ranks."""
    res = {}
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    big = [("twoshot_tma", lib.ALGO_TWOSHOT_TMA), ("twoshot", lib.ALGO_TWOSHOT)]
    if comm.has_multicast:
        big.insert(0, ("nvls", lib.ALGO_NVLS))
    cases = [("f32", torch.float32, lib.F32, 4), ("bf16", torch.bfloat16, lib.BF16, 2)]
    half = (comm.arena_bytes // 2) // 4096 * 4096

    def rel(a, b):
        a, b = a.double(), b.double()
        d = float(b.abs().max())
        return float((a - b).abs().max()) / (d if d > 0 else 1.0)

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def ordered_sum(mine, tdt):
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        acc = parts[0].float()
        for p in parts[1:]:
            acc = acc + p.float()
        return acc.to(tdt)

    for dname, tdt, code, es in cases:
        epv = 16 // es
        for n_el in (epv * 1001, (1 << 20) // es + epv * 3, (6 << 20) // es + epv * 7):
            nbytes = n_el * es
            gen = torch.Generator().manual_seed(77 + 13 * rank + n_el % 1009)
            mine = ((torch.rand(n_el, generator=gen) * 2 - 1)).to(tdt).to(dev)
            want_nccl = mine.clone()
            dist.all_reduce(want_nccl)
            want_ord = ordered_sum(mine, tdt)
            for aname, algo in big:
                comm.view(0, nbytes, tdt).copy_(mine)
                torch.cuda.synchronize()
                dist.barrier()
                comm.allreduce(0, 0, nbytes, code, algo, 0, sp)
                stream.synchronize()
                comm.poll_error()
                got = comm.view(0, nbytes, tdt).clone()
                key = f"{aname}/{dname}"
                cur = res.setdefault(key, {"vs_nccl": 0.0, "vs_ordered": 0.0})
                cur["vs_nccl"] = max(cur["vs_nccl"], rel(got, want_nccl))
                cur["vs_ordered"] = max(cur["vs_ordered"], rel(got, want_ord))
                dist.barrier()
        # the fused single-launch kernels: LL (<= 32 KiB) and one-barrier (<= 256 KiB), multi-tensor tables
        for fname, lens in (("fused_ll", [20, 1, 513, 7]), ("fused_1barrier", [30000, 5, 4099, 12001])):
            gen = torch.Generator().manual_seed(5 + rank)
            grads = [((torch.rand(n, generator=gen) * 2 - 1)).to(tdt).to(dev) for n in lens]
            flat_in = torch.cat(grads)
            scaled = (flat_in.float() * (1.0 / world)).to(tdt)  # the kernel pre-scales like the reference
            want_nccl = scaled.clone()
            dist.all_reduce(want_nccl)
            want_ord = ordered_sum(scaled, tdt)
            offs, off = [], 0
            for n in lens:
                offs.append(off)
                off += (n + epv - 1) // epv * epv
            bucket_bytes = off * es
            region = comm.fused_region_bytes(bucket_bytes)
            comm.clear(half, (region + 15) // 16 * 16, sp)
            stream.synchronize()
            dist.barrier()
            comm.fused_bucket(TensorTable([g.data_ptr() for g in grads], offs, lens), code, half, bucket_bytes,
                              1.0 / world, sp)
            stream.synchronize()
            comm.poll_error()
            got = torch.cat(grads)
            key = f"{fname}/{dname}"
            res[key] = {"vs_nccl": rel(got, want_nccl), "vs_ordered": rel(got, want_ord)}
            dist.barrier()
    out = {}
    for key, v in res.items():
        out[key] = {k: reduce_max(x) for k, x in v.items()}
    tol = {"f32": 1e-6, "bf16": 2.0 ** -6}
    ok = all(v["vs_nccl"] <= tol[k.split("/")[1]] for k, v in out.items())
    # P2P kernels sum in rank order in fp32 and round once: exactly the ordered sum
    exact = all(v["vs_ordered"] == 0.0 for k, v in out.items() if not k.startswith("nvls"))
    return {"max_rel": out, "tolerance_vs_nccl": tol, "ok": bool(ok), "p2p_bit_exact_vs_rank_ordered_sum": bool(exact),
            "sizes_elements": "1001 vectors, 1 MiB + 3 vectors, 6 MiB + 7 vectors; fused: 4-tensor ragged tables",
            "metric": "max|a-b| / max|b| per buffer, max over sizes and ranks"}


def nvlink_block(torch, dist, comm, world, rank, local, dev, lib):
    """(1) peer-read bandwidth: every rank pulls 256 MiB from its right neighbour with 16-byte loads, all ranks at
    once — the per-direction NVLink rate a kernel reaches on this box; (2) NVML NVLink DATA byte counters around a
    known number of allreduce launches: measured wire bytes per launch against the algorithm's (1+1/ws) S (NVLS) or
    2(ws-1)/ws S (two-shot)."""
    out = {}
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    S = min(256 << 20, (comm.arena_bytes // 2) // (1 << 20) * (1 << 20))
    peer = (rank + 1) % world
    best = None
    for ctas in (0, 64, 148):
        torch.cuda.synchronize()
        dist.barrier()
        ts = []
        for rep in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            comm.peer_read(peer, 0, S, S, sp, ctas)
            b.record(stream)
            stream.synchronize()
            if rep >= 1:
                ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gbs = S / (float(t.item()) * 1e-3) / 1e9
        if best is None or gbs > best["gbs"]:
            best = {"gbs": gbs, "ctas": ctas or "default"}
    out["peer_read_gbs"] = best["gbs"]
    out["peer_read"] = {"bytes": S, "ctas": best["ctas"],
                        "how": "b200ddp_peer_read: 16-byte loads from the right neighbour's arena, all ranks "
                               "simultaneously, CUDA events, median of 4, max over ranks"}
    # ---- NVML counters ----
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        TX, RX = 138, 139  # NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX, KiB, scopeId UINT_MAX = all links

        def read():
            vals = pynvml.nvmlDeviceGetFieldValues(handle, [(TX, 0xFFFFFFFF), (RX, 0xFFFFFFFF)])
            got = []
            for v in vals:
                if v.nvmlReturn != 0:
                    raise RuntimeError(f"NVML field {v.fieldId}: return code {v.nvmlReturn}")
                got.append(int(v.value.ullVal) * 1024)
            return got

        Sb = 64 << 20
        launches = 40
        counters = {}
        algos = [("twoshot_tma", lib.ALGO_TWOSHOT_TMA, 2 * (world - 1) / world)]
        if comm.has_multicast:
            algos.insert(0, ("nvls", lib.ALGO_NVLS, 1 + 1 / world))
        for name, algo, factor in algos:
            torch.cuda.synchronize()
            dist.barrier()
            time.sleep(0.25)
            t0, r0 = read()
            for _ in range(launches):
                comm.allreduce(0, 0, Sb, lib.F32, algo, 0, sp)
            stream.synchronize()
            comm.poll_error()
            dist.barrier()
            time.sleep(0.25)
            t1, r1 = read()
            counters[name] = {"bucket_bytes": Sb, "launches": launches, "tx_bytes_per_launch": (t1 - t0) / launches,
                              "rx_bytes_per_launch": (r1 - r0) / launches,
                              "expected_per_direction": factor * Sb,
                              "tx_over_expected": (t1 - t0) / launches / (factor * Sb)}
        out["nvml_counters"] = counters
        out["nvml_counters_note"] = ("NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX summed over the GPU's links, rank 0; "
                                     "expected = (1+1/ws) S for NVLS (S pulled by the switch + S/ws multicast store), "
                                     "2(ws-1)/ws S for the two-shot")
        pynvml.nvmlShutdown()
    except Exception as exc:  # counters are evidence, not a dependency
        out["nvml_counters"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def stock_ddp_leg(torch, h: Harness, wl, workload, bsz, local, rank, devb, steps, warmup, blocks, cap_mb, views):
    """Stock torch-DDP over NCCL: same model, loop, batches and timing (the same-hardware comparator)."""
    trainer = build_trainer(torch, wl, workload, bsz, local, rank, True, cap_mb, views, "off")
    nb = len(devb)
    for i in range(max(warmup, 4)):
        trainer._run_batch(*devb[i % nb])
    ms, all_ms = h.timed_blocks(lambda i: trainer._run_batch(*devb[i % nb]), steps, blocks)
    out = {"value": bsz * h.world * steps / (ms / 1e3), "ms_per_step": ms / steps, **block_stats(all_ms, steps),
           "what": "torch.nn.parallel.DistributedDataParallel + NCCL, eager, device-resident batches"}
    del trainer
    return out


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
    blocks = blocks_for(args.workload, steps, args.blocks)

    multigpu_torchrun.ddp_setup()  # init_process_group("nccl") + bind to LOCAL_RANK's GPU
    dev = torch.device("cuda", local)
    h = Harness(torch, dist, dev, world, rank)
    wl = workloads.get(args.workload)
    bsz = wl.per_gpu_batch
    stock = args.impl == "nccl"
    legs_failed = {}

    def guarded(name, fn, default=None):
        try:
            return fn()
        except Exception as exc:  # an optional leg must never cost the line
            legs_failed[name] = f"{type(exc).__name__}: {exc}"
            return default

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    host, devb, shard = make_batches(torch, wl, args.workload, bsz, local, rank)
    if args.fuse_sgd and not stock:
        os.environ["B200DDP_FUSE_SGD"] = "1"
    trainer = build_trainer(torch, wl, args.workload, bsz, local, rank, stock, args.bucket_cap_mb, args.views, args.graph)
    os.environ.pop("B200DDP_FUSE_SGD", None)  # the comparator and ResNet-50 legs stay on the default path
    ddp = trainer.model

    class _NoComm:  # stock DDP has no b200ddp communicator
        launch_count = 0
        mem_mode = "nccl"
        has_multicast = None

    comm = _NoComm() if stock else ddp.engine.comm

    # ---- eager record (B200DDP_GRAPH=0 behaviour) on the same trainer, before anything is captured -----------
    clocks = ClockSampler(local).start() if rank == 0 else None
    st = trainer._state()
    mode_wanted = st["mode"]
    eager = None
    if not stock and mode_wanted != "off":
        st["mode"] = "off"
        for i in range(warmup):
            trainer._run_batch(*devb[i % len(devb)])
        ems, eall = h.timed_blocks(lambda i: trainer._run_batch(*devb[i % len(devb)]), steps, min(blocks, 5))
        eager = {"value": bsz * world * steps / (ems / 1e3), "ms_per_step": ems / steps, **block_stats(eall, steps),
                 "note": "B200DDP_GRAPH=0: the same step with every kernel launched eagerly (hook loop: "
                         + ("C++" if getattr(ddp.reducer, "_native", None) is not None else "Python") + ")"}
        st["mode"] = mode_wanted
        st["eager_run"] = max(st["eager_run"], 3)

    # ---- value + e2e -----------------------------------------------------------------------------------------
    h.barrier_sync()
    if clocks:
        clocks.mark()
    value, e2e = time_training(h, trainer, host, devb, shard, steps, warmup, blocks, bsz, resident_ok=not stock)
    if clocks:
        clocks.mark()
    clock_info = clocks.stop() if clocks else None
    replayed = (not stock) and ("graph" in value["mode"])
    launches = (value["launches_per_replay"] * steps * blocks) if replayed else None

    if stock:
        if rank == 0:
            emit({"impl": "nccl", "metric": "samples/sec", "value": value["value"], "unit": "samples/s", "n_gpus": world,
                  "steps": steps, "warmup": warmup, "ms_per_step": value["ms_per_step"], "higher_is_better": True,
                  "scaling": "weak", "data": "synthetic", "e2e": e2e, "clocks": clock_info,
                  "timed_blocks": value["timed_blocks"], "config": bench_config(args.workload, world, bsz),
                  "impl_config": {"stack": "stock torch-DDP + NCCL on the same GPUs (comparator)",
                                  "gradient_as_bucket_view": bool(args.views)}})
        dist.destroy_process_group()
        return 0

    if launches is None:  # eager: count them over one more block
        n0 = comm.launch_count
        for i in range(steps):
            trainer._run_batch(*devb[i % len(devb)])
        torch.cuda.synchronize()
        launches = (comm.launch_count - n0) * blocks

    # ---- nvlink peak (N>1) ---------------------------------------------------------------------------------------
    nvl_peak, nvl_src = NVLINK_NOMINAL, (f"nominal NVLink 5: {NVLINK_NOMINAL:.0f} GB/s per direction per GPU (pool guide: "
                                         f"measured peer copy {NVLINK_POOL_MEASURED:.0f} GB/s; this run's own peer-read "
                                         "rate is in nvlink_peak_measured)")
    nvlink = None
    side_comm = None
    if world > 1:
        side_comm = guarded("side_comm", lambda: P2PComm.from_process_group((512 << 20) + (8 << 20), local))
    if side_comm is not None:
        nvlink = guarded("nvlink", lambda: nvlink_block(torch, dist, side_comm, world, rank, local, dev, _lib))
        if nvlink and nvlink.get("peer_read_gbs"):
            nvlink["nominal_gbs"] = NVLINK_NOMINAL

    # ---- roofline of the step's own largest bucket ---------------------------------------------------------------
    roofline_step = guarded("roofline_step", lambda: bucket_roofline(torch, h, ddp, hbm_peak, peak_src, nvl_peak, nvl_src))

    # ---- parity (N>1) ----------------------------------------------------------------------------------------
    parity = None
    if side_comm is not None:
        parity = guarded("parity", lambda: parity_block(torch, dist, side_comm, world, rank, dev, _lib, TensorTable))
        guarded("side_comm_destroy", lambda: (torch.cuda.synchronize(), dist.barrier(), side_comm.destroy()))

    # ---- bucket-size sweep ------------------------------------------------------------------------------
    sweep = []
    if not args.no_sweep:
        sweep = guarded("bucket_sweep", lambda: bucket_sweep(
            torch, dist, comm_factory=lambda nbytes: P2PComm.from_process_group(nbytes, local), world=world, rank=rank,
            dev=dev, hbm_peak=hbm_peak, nvl_peak=nvl_peak, max_mib=args.sweep_max_mib, lib=_lib,
            TensorTable=TensorTable), default=[])

    # ---- the same-hardware comparator: stock DDP over NCCL --------------------------------------------------------
    nccl_ddp = None
    if not args.no_nccl:
        nccl_ddp = guarded("nccl_ddp", lambda: stock_ddp_leg(torch, h, wl, args.workload, bsz, local, rank, devb, steps,
                                                             warmup, min(blocks, 5), args.bucket_cap_mb, args.views))

    # ---- extra: ResNet-50 (configs[3]) beside every toy line --------------------------------------------------
    extra = None
    if args.workload == "toy" and not args.no_extra:
        extra = guarded("extra.resnet50", lambda: resnet_leg(torch, h, local, rank, args.graph, args.no_nccl,
                                                             hbm_peak, peak_src, nvl_peak, nvl_src))

    # ---- cpu baseline (N=1, rank 0) ----------------------------------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        def run_cpu():
            s, w = reference_budget(args.workload, None, None)
            r = reference_run(args.workload, 1, s, w)
            return {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                    "sample": f"{s} steps (batch {bsz}) after {w} warm-up of stock torch-DDP on gloo/CPU, 1 process x "
                              f"{r['threads_per_rank']} threads, via oracle/ref_ddp.py"}
        cpu = guarded("cpu_baseline", run_cpu)

    line = {
        "impl": "b200", "metric": "samples/sec", "value": value["value"], "unit": "samples/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": value["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if wl.dtype == torch.bfloat16 else "f32",
        "data": "synthetic", "config": bench_config(args.workload, world, bsz),
        "timed_blocks": value["timed_blocks"], "block_ms": {k: value[k] for k in ("block_ms_min", "block_ms_median", "block_ms_max")},
        "value_mode": value["mode"],
        "impl_config": {"bucket_bytes": ddp.reducer.bucket_sizes(),
                        "bucket_algos": [("fused" if b.fused else b.algo) for b in ddp.reducer.buckets],
                        "mem_mode": comm.mem_mode, "multicast": comm.has_multicast, "cuda_graph": replayed,
                        "graph_mode": mode_wanted, "arena_bytes": comm.arena_bytes,
                        "gradient_as_bucket_view": bool(ddp.gradient_as_bucket_view),
                        "fused_sgd": getattr(ddp, "_fused_optimizer", None) is not None,
                        "hook_loop": "C++" if getattr(ddp.reducer, "_native", None) is not None else "Python"},
        "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline_step": roofline_step,
        "bucket_sweep": sweep, "cpu_baseline": cpu, "eager": eager, "nccl_ddp": nccl_ddp, "parity": parity,
        "nvlink": nvlink, "extra": extra,
    }
    if nvlink and nvlink.get("peer_read_gbs"):
        line["nvlink_peak_measured"] = {"value": nvlink["peer_read_gbs"], "unit": "GB/s per direction per GPU",
                                        "peak_source": "measured: b200ddp_peer_read in this run"}
    # `roofline`: the dominant bandwidth-bound kernel at the bucket size DDP actually produces (25 MiB cap):
    # N=1 flatten of the ResNet-50-shaped bucket against HBM; N>1 the AUTO allreduce at 25 MiB against NVLink.
    roofline = None
    if sweep:
        if world == 1:
            cand = [p for p in sweep if p.get("kernel") == "flatten_scale[resnet50 bucket]" and "achieved" in p] or \
                   [p for p in sweep if p.get("kernel") == "flatten_scale" and p["bytes"] == (25 << 20)] or \
                   [p for p in sweep if p.get("kernel") == "flatten_scale"]
        else:
            ours = [p for p in sweep if p.get("auto")]
            cand = [p for p in ours if p["bytes"] == (25 << 20)] or ours[-1:]
        if cand:
            best = cand[-1]
            traffic = None
            try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                traffic = tr.get(f"{best['kernel']}@{best['bytes']}", {}).get("dram_bytes_per_launch")
            except OSError:
                pass
            roofline = {"bound": best["bound"], "achieved": best["achieved"], "peak": best["peak"], "unit": "GB/s",
                        "frac": best["frac"], "traffic": traffic, "kernel": best["kernel"], "bucket_bytes": best["bytes"],
                        "avg_launch_ms": best["ms"],
                        "algorithmic_bytes_per_launch": (2 * best["bytes"] if world == 1 else
                                                         best["bytes"] * 2 * (world - 1) / world),
                        "peak_source": peak_src if world == 1 else nvl_src,
                        "where": "bucket_sweep of this run (CUDA-graph replay timed with CUDA events on the replaying "
                                 "stream); the timed step's own bucket is in roofline_step"}
            if world > 1 and nvlink and nvlink.get("peer_read_gbs"):
                roofline["frac_of_measured_peer_read"] = best["achieved"] / nvlink["peer_read_gbs"]
            if world > 1 and nvlink:
                # no DRAM counters for kernels that wait for a peer (ncu cannot run them): the measured quantity is
                # the NVLink wire traffic — NVML DATA TX bytes per launch of this algorithm (64 MiB x 40 launches in
                # this run), scaled to this bucket size
                name = best["kernel"].replace("allreduce_", "")
                ctr = (nvlink.get("nvml_counters") or {}).get(name)
                if isinstance(ctr, dict) and ctr.get("tx_over_expected"):
                    factor = (1 + 1 / world) if name == "nvls" else 2 * (world - 1) / world
                    roofline["traffic"] = ctr["tx_over_expected"] * factor * best["bytes"]
                    roofline["traffic_kind"] = ("NVLink bytes sent per launch and GPU (NVML DATA TX counters, measured / "
                                                f"algorithmic = {ctr['tx_over_expected']:.5f}); wire bytes = "
                                                + ("(1+1/ws) S" if name == "nvls" else "2(ws-1)/ws S"))
    line["roofline"] = roofline if roofline is not None else roofline_step
    if legs_failed:
        line["legs_failed"] = legs_failed
    if rank == 0:
        emit(line)
    multigpu_torchrun.destroy_process_group()
    return 0


def resnet_leg(torch, h: Harness, local, rank, graph_mode, no_nccl, hbm_peak, peak_src, nvl_peak, nvl_src):
    """configs[3] beside the toy line: ResNet-50, batch 32/GPU, ~100 MB of gradients in 5 buckets, through the same
    Trainer surface (default DDP(model, device_ids=[i]) call, Trainer's default graph mode)."""
    from b200ddp import workloads

    wl = workloads.get("resnet50")
    bsz = wl.per_gpu_batch
    steps, warmup, blocks = 8, 4, 3
    host, devb, _ = make_batches(torch, wl, "resnet50", bsz, local, rank)
    trainer = build_trainer(torch, wl, "resnet50", bsz, local, rank, False, None, False, graph_mode)
    value, e2e = time_training(h, trainer, host[:4], devb[:4], None, steps, warmup, blocks, bsz, resident_ok=False)
    out = {"samples_s": value["value"], "ms_per_step": value["ms_per_step"], "steps": steps,
           "timed_blocks": value["timed_blocks"], "mode": value["mode"], "e2e_samples_s": e2e["value"],
           "e2e_ms_per_step": e2e["ms_per_step"], "bucket_bytes": trainer.model.reducer.bucket_sizes(),
           "workload": WORKLOAD_DESC["resnet50"]}
    try:
        out["roofline_step"] = bucket_roofline(torch, h, trainer.model, hbm_peak, peak_src, nvl_peak, nvl_src)
    except Exception as exc:
        out["roofline_step"] = {"error": f"{type(exc).__name__}: {exc}"}
    trainer._graph = None
    trainer.model.shutdown()
    del trainer
    torch.cuda.empty_cache()
    if not no_nccl:
        try:
            out["nccl_ddp"] = stock_ddp_leg(torch, h, wl, "resnet50", bsz, local, rank, devb[:4], steps, warmup, blocks,
                                            None, False)
        except Exception as exc:
            out["nccl_ddp"] = {"error": f"{type(exc).__name__}: {exc}"}
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
