"""Grid-size / algorithm sweep of the allreduce kernels on a real multi-GPU world (torchrun).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \\
        tools/tune_allreduce.py [sizes_mib=1,16,64,256] [ctas=0,16,32,64,128] [dtype=f32]

Each (algorithm, size, grid cap) is timed as a CUDA graph of back-to-back launches rotating through
> L2 of distinct arena (see bench.py::bucket_sweep).  Rank 0 prints one JSON line per point and a table.
"""
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ddp  # noqa: E402,F401
from b200ddp import _lib, destroy_process_group, init_process_group  # noqa: E402
from b200ddp.comm import P2PComm  # noqa: E402


def main():
    kw = dict(a.split("=", 1) for a in sys.argv[1:])
    sizes = [int(float(x) * (1 << 20)) for x in kw.get("sizes_mib", "0.25,1,4,16,64,256").split(",")]
    ctas_list = [int(x) for x in kw.get("ctas", "0,16,32,64,128").split(",")]
    dt = kw.get("dtype", "f32")
    code, tdt = {"f32": (_lib.F32, torch.float32), "bf16": (_lib.BF16, torch.bfloat16)}[dt]
    init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    biggest = max(sizes)
    slots = max(2, (288 << 20) // biggest + 1)
    comm = P2PComm.from_process_group(biggest * slots + (8 << 20), dev.index)
    comm.arena().view(torch.float32).uniform_(-1e-3, 1e-3)
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    algos = [("oneshot", 1), ("twoshot", 2), ("tma", 3)] + ([("nvls", 4)] if comm.has_multicast else [])
    if "algos" in kw:
        algos = [a for a in algos if a[0] in kw["algos"].split(",")]
    rows = []

    def time_graph(fn, iters):
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(iters):
                fn(i)
        ts = []
        with torch.cuda.stream(stream):
            for rep in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                g.replay()
                b.record(stream)
                stream.synchronize()
                if rep >= 2:
                    ts.append(a.elapsed_time(b) / iters)
        t = torch.tensor([statistics.median(ts)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for S in sizes:
        reps = max(2, (288 << 20) // S + 1)
        iters = max(40 if S <= (4 << 20) else (20 if S <= (64 << 20) else 8), reps - 1)
        iters = min(iters, 400)
        for name, algo in algos:
            if name == "oneshot" and S > (4 << 20):
                continue
            for ctas in ctas_list:
                def fn(i):
                    off = (i % (reps - 1)) * S if name != "oneshot" else 0
                    comm.allreduce(off, off if name != "oneshot" else (reps - 1) * S, S, code, algo, ctas, sp)
                ms = time_graph(fn, iters)
                comm.poll_error()
                bus = S * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
                row = dict(algo=name, mib=S / (1 << 20), ctas=ctas, us=ms * 1e3, bus_gbs=bus, algbw_gbs=S / (ms * 1e-3) / 1e9,
                           world=world, dtype=dt)
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
    # the fused single-launch path (flatten+scale -> allreduce -> unflatten): LL protocol <= 32 KiB, one barrier above
    from b200ddp.comm import TensorTable

    for kib in [float(x) for x in kw.get("fused_kib", "0.09375,4,16,32,64,256").split(",") if x]:
        n_el = max(4, int(kib * 1024) // 4 // 4 * 4)
        nbytes = n_el * 4
        g = torch.rand(n_el, device=dev)
        table = TensorTable([g.data_ptr()], [0], [n_el])
        region = comm.fused_region_bytes(nbytes)
        comm.clear(0, region, sp)
        stream.synchronize()
        dist.barrier()
        ms = time_graph(lambda i: comm.fused_bucket(table, code if dt == "f32" else _lib.F32, 0, nbytes, 1.0 / world, sp), 50)
        comm.poll_error()
        row = dict(algo="fused_ll" if nbytes <= 32768 else "fused_1barrier", mib=nbytes / (1 << 20), ctas=0, us=ms * 1e3,
                   bus_gbs=nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9, world=world, dtype="f32")
        if rank == 0:
            print(json.dumps(row), flush=True)
    comm.arena().view(torch.float32).uniform_(-1e-3, 1e-3)
    torch.cuda.synchronize()
    dist.barrier()
    # NCCL on the same box, same timing method
    for S in (sizes if kw.get("nccl", "1") == "1" else []):
        reps = max(2, min(8, (288 << 20) // S + 1))
        bufs = [torch.empty(S // 4, device=dev).uniform_(-1e-3, 1e-3) for _ in range(reps)]
        with torch.cuda.stream(stream):
            for b in bufs[:2]:
                dist.all_reduce(b)
        iters = 40 if S <= (4 << 20) else (20 if S <= (64 << 20) else 8)
        try:
            ms = time_graph(lambda i: dist.all_reduce(bufs[i % reps]), iters)
            how = "graph"
        except Exception as exc:
            how = f"eager:{type(exc).__name__}"
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                dist.all_reduce(bufs[i % reps])
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
        row = dict(algo="nccl", mib=S / (1 << 20), ctas=-1, us=ms * 1e3, bus_gbs=S * 2 * (world - 1) / world / (ms * 1e-3) / 1e9,
                   world=world, how=how)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del bufs
    torch.cuda.synchronize()
    dist.barrier()
    comm.destroy()
    destroy_process_group()


if __name__ == "__main__":
    main()
