"""Debug aid: the DDP constructor sequence of a 2-rank loop-back world, with variants."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant = sys.argv[1] if len(sys.argv) > 1 else "A"
if variant == "D":
    os.environ["CUDA_MODULE_LOADING"] = "EAGER"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

import b200ddp  # noqa
from b200ddp import _lib
from b200ddp.comm import P2PComm, TensorTable

torch.cuda.init()
comms = P2PComm.loopback_world(2, 0, 8 << 20)
for c in comms:
    c.set_timeout_ms(1500)
prio = 0 if variant == "C" else -1
cs = [torch.cuda.Stream(priority=prio) for _ in range(2)]
us = [torch.cuda.Stream() for _ in range(2)]
w = [torch.rand(21, device="cuda") for _ in range(2)]
torch.cuda.synchronize()
if variant == "E":  # pre-load every kernel once on a 1-rank context
    c1 = P2PComm(0, 1, 0, 1 << 20, _lib.MEM_IPC)
    t = TensorTable([w[0].data_ptr()], [0], [21])
    c1.flatten_scale(t, _lib.F32, 0, 1.0, us[0].cuda_stream)
    c1.unflatten(t, _lib.F32, 0, us[0].cuda_stream)
    torch.cuda.synchronize()
t0 = time.time()


def sw(c, r, waiter, signaled):
    """waiter waits for signaled; variants choose the mechanism"""
    if variant in "FG" and variant == "F":
        return
    if variant == "I":
        e = torch.cuda.Event()
        e.record(signaled)
        waiter.wait_event(e)
        return
    c.stream_wait(waiter.cuda_stream, signaled.cuda_stream)


for r in range(2):
    c, s, u = comms[r], cs[r].cuda_stream, us[r].cuda_stream
    table = TensorTable([w[r].data_ptr()], [0], [21])
    sw(c, r, cs[r], us[r])
    if r == 0 and variant != "H":
        c.flatten_scale(table, _lib.F32, 0, 1.0, s)
    c.broadcast(0, 96, 0, s, 0)
    if r != 0 and variant != "H":
        c.unflatten(table, _lib.F32, 0, s)
    if variant != "G":
        sw(c, r, us[r], cs[r])
        sw(c, r, cs[r], us[r])
    if variant not in "BH":
        c.clear(1 << 20, 4096, s)
    c.barrier(s)
    print(f"[{variant}] rank {r} enqueued at {time.time()-t0:.3f}s", flush=True)
torch.cuda.synchronize()
print(f"[{variant}] synchronized after {time.time()-t0:.3f}s", flush=True)
for c in comms:
    try:
        c.poll_error()
        print(f"[{variant}] rank {c.rank} ok")
    except RuntimeError as e:
        print(f"[{variant}] rank {c.rank} ERROR {e}")
print(torch.equal(w[0], w[1]))
