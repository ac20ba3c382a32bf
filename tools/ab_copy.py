"""A/B of the flatten kernel's work split on ONE box: even split vs round-robin tiles, interleaved."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ddp  # noqa
from b200ddp import _lib
from b200ddp.comm import P2PComm, TensorTable

S_list = [16 << 20, 25 << 20, 64 << 20, 256 << 20]
comm = P2PComm(0, 1, 0, (256 << 20) * 2 + (8 << 20), _lib.MEM_AUTO)
stream = torch.cuda.Stream()
sp = stream.cuda_stream
for S in S_list:
    n_t, per = 8, S // 4 // 8
    reps = max(2, (288 << 20) // S + 1)
    srcs = [[torch.rand(per, device="cuda") for _ in range(n_t)] for _ in range(reps)]
    offs = [i * per for i in range(n_t)]
    tables = [TensorTable([t.data_ptr() for t in ts], offs, [per] * n_t) for ts in srcs]
    iters = max(8, reps)
    res = {"even": [], "tile": []}
    graphs = {}
    for mode in ("even", "tile"):
        os.environ["B200DDP_COPY_SPLIT"] = mode  # read at launch time
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(iters):
                comm.flatten_scale(tables[i % reps], _lib.F32, (i % (reps - 1)) * S if reps > 2 else 0, 0.5, sp)
        graphs[mode] = g
    with torch.cuda.stream(stream):
        for rep in range(8):
            for mode in ("even", "tile"):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); graphs[mode].replay(); b.record(stream); stream.synchronize()
                if rep >= 2:
                    res[mode].append(a.elapsed_time(b) / iters)
    for mode in res:
        ms = statistics.median(res[mode])
        print(f"S={S>>20:4d} MiB {mode}: {ms*1e3:7.2f} us  {2*S/(ms*1e-3)/1e9:7.1f} GB/s", flush=True)
    del srcs, tables, graphs
comm.destroy()
