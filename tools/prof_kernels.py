"""Single-GPU launches of the HBM-bound kernels for ncu (one GPU only: ncu replays every kernel in
isolation, so kernels that wait for a peer must never run under it).

    ncu --set full --clock-control none --import-source on -k regex:multi_tensor_copy -s 4 -c 4 \\
        -o gpurun_out/prof_flatten python tools/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ddp  # noqa: E402,F401
from b200ddp import _lib  # noqa: E402
from b200ddp.comm import P2PComm, TensorTable  # noqa: E402

S = int(os.environ.get("PROF_BYTES", str(64 << 20)))
comm = P2PComm(0, 1, 0, S + (1 << 20), _lib.MEM_AUTO)
stream = torch.cuda.Stream()
n_t = 8
per = S // 4 // n_t
sets = [[torch.rand(per, device="cuda") for _ in range(n_t)] for _ in range(4)]  # 256 MiB > L2
offs = [i * per for i in range(n_t)]
tables = [TensorTable([t.data_ptr() for t in ts], offs, [per] * n_t) for ts in sets]
torch.cuda.synchronize()
for i in range(4):
    comm.flatten_scale(tables[i], _lib.F32, 0, 0.125, stream.cuda_stream)
    comm.unflatten(tables[(i + 1) % 4], _lib.F32, 0, stream.cuda_stream)
# the largest real ResNet-50 bucket (bench.resnet_table: two 9 MiB convolutions + dozens of 256 B - 8 KiB
# BatchNorm vectors): boundary tiles and straight tiles in one launch
import bench  # noqa: E402

S2, n_t2, tables2, keep2 = bench.resnet_table(torch, TensorTable, torch.device("cuda", 0), 4)
comm2 = P2PComm(0, 1, 0, S2 + (1 << 20), _lib.MEM_AUTO)
for i in range(4):
    comm2.flatten_scale(tables2[i], _lib.F32, 0, 0.125, stream.cuda_stream)
    comm2.unflatten(tables2[(i + 1) % 4], _lib.F32, 0, stream.cuda_stream)
torch.cuda.synchronize()
comm2.destroy()
# the toy model's 84-byte bucket through the fused kernel (world 1)
g = [torch.rand(20, device="cuda"), torch.rand(1, device="cuda")]
t3 = TensorTable([x.data_ptr() for x in g], [0, 20], [20, 1])
for _ in range(4):
    comm.fused_bucket(t3, _lib.F32, 0, 96, 1.0, stream.cuda_stream)
torch.cuda.synchronize()
comm.destroy()
print("prof_kernels done")
