"""A/B of the flatten/unflatten kernel's tile size on ONE box, interleaved: B200DDP_COPY_UNROLL = 2 | 4 | 8
(vectors in flight per thread) on clean 8-tensor buckets of the sizes DDP produces and on the largest
ResNet-50 bucket (two 9 MiB convolutions + dozens of 256 B - 8 KiB BatchNorm vectors).  CUDA-graph replays timed
with CUDA events; every graph streams > L2 of distinct buffers."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ddp  # noqa
from b200ddp import _lib
from b200ddp.comm import P2PComm, TensorTable
import bench

PEAK = 6551.0
comm = P2PComm(0, 1, 0, (576 << 20) + (8 << 20), _lib.MEM_AUTO)
stream = torch.cuda.Stream()
sp = stream.cuda_stream
dev = torch.device("cuda", 0)


VARIANTS = [("4", "1"), ("4", "2"), ("4", "4"), ("4", "64"), ("2", "64"), ("8", "1")]  # (COPY_UNROLL, COPY_WAVES)


def measure(label, S, tables, reps, slot):
    iters = max(16, reps)
    graphs = {}
    for u, w in VARIANTS:
        os.environ["B200DDP_COPY_UNROLL"] = u  # both read at launch time
        os.environ["B200DDP_COPY_WAVES"] = w
        for kind in ("flatten", "unflatten"):
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(iters):
                    off = (i % (reps - 1)) * slot
                    if kind == "flatten":
                        comm.flatten_scale(tables[i % reps], _lib.F32, off, 0.5, sp)
                    else:
                        comm.unflatten(tables[i % reps], _lib.F32, off, sp)
            graphs[(u, w, kind)] = g
    # comparator: torch's own copy of one contiguous tensor of the same size
    a = [torch.rand(S // 4, device=dev) for _ in range(reps)]
    b = [torch.empty(S // 4, device=dev) for _ in range(reps)]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(iters):
            b[i % reps].copy_(a[i % reps])
    graphs[("torch", "-", "copy_")] = g
    res = {k: [] for k in graphs}
    with torch.cuda.stream(stream):
        for rep in range(9):
            for k, g in graphs.items():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); g.replay(); e1.record(stream); stream.synchronize()
                if rep >= 2:
                    res[k].append(e0.elapsed_time(e1) / iters)
    for (u, w, kind), ts in res.items():
        ms = statistics.median(ts)
        gbs = 2 * S / (ms * 1e-3) / 1e9
        print(f"{label:28s} S={S/2**20:6.2f} MiB unroll={u:>5s} waves={w:>2s} {kind:9s}: {ms*1e3:7.2f} us  {gbs:7.1f} GB/s  "
              f"{gbs/PEAK:.3f} of {PEAK:.0f}", flush=True)
    del a, b


for S in (8 << 20, 25 << 20, 64 << 20, 256 << 20):
    n_t, per = 8, S // 4 // 8
    reps = max(2, (288 << 20) // S + 1)
    srcs = [[torch.rand(per, device=dev) for _ in range(n_t)] for _ in range(reps)]
    tables = [TensorTable([t.data_ptr() for t in ts], [i * per for i in range(n_t)], [per] * n_t) for ts in srcs]
    measure("clean 8 tensors", S, tables, reps, S)
    del srcs, tables
reps = 10
S, n_t, tables, keep = bench.resnet_table(torch, TensorTable, dev, reps)
measure(f"resnet50 bucket ({n_t} tensors)", S, tables, reps, (S + 4095) // 4096 * 4096)
# the bucket with the MOST tiny tensors: ResNet-50's last-ready bucket (conv1/bn1/layer1: ~50 tensors in 8 MiB)
from torchvision.models import resnet50
from b200ddp.bucketing import compute_bucket_assignment_by_size, layout_bucket
with torch.device("meta"):
    m = resnet50()
ps = list(m.parameters())
sizes = [p.numel() * 4 for p in ps]
idx, _ = compute_bucket_assignment_by_size(sizes, [0] * len(ps), [1 << 20, 25 << 20])
many = max(idx, key=len)
numels = [ps[i].numel() for i in many]
spec = layout_bucket(0, many, numels, 0, 4)
keep2, tables2 = [], []
reps2 = max(2, (288 << 20) // spec.nbytes + 1)
for _ in range(reps2):
    ts = [torch.rand(n, device=dev) for n in numels]
    keep2.append(ts)
    tables2.append(TensorTable([t.data_ptr() for t in ts], spec.offsets_el, spec.lengths_el))
measure(f"resnet50 bucket ({len(numels)} tensors)", spec.nbytes, tables2, reps2, (spec.nbytes + 4095) // 4096 * 4096)
os.environ.pop("B200DDP_COPY_UNROLL", None)
os.environ.pop("B200DDP_COPY_WAVES", None)
comm.destroy()
