"""Where does a replayed toy step spend its time?  Host cost of every call of the pipelined host path
(perf_counter around each, 3000 iterations) and device time of each captured step kind (events around a replay,
device idle before).  One GPU, world 1."""
import os, sys, time, statistics, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ddp
from b200ddp import multigpu, utils
from b200ddp.utils import DeviceShardLoader, MyTrainDataset

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29677")
multigpu.ddp_setup(0, 1)
multigpu.Trainer.loss_fn = staticmethod(torch.nn.MSELoss())
torch.manual_seed(0)
model = torch.nn.Linear(20, 1)
opt = torch.optim.SGD(model.parameters(), lr=1e-3)
t = multigpu.Trainer(model, [], opt, 0, save_every=10**9)
ds = MyTrainDataset(32768, seed=1)
host = [(ds.x[i:i + 32].pin_memory(), ds.y[i:i + 32].pin_memory()) for i in range(0, 32 * 64, 32)]
dev = [(x.cuda(), y.cuda()) for x, y in host]
for i in range(6):
    t._run_batch(*dev[i])
for i in range(8):
    out = t._run_batch(*host[i])
out.item()
from torch.utils.data.distributed import DistributedSampler
shard = DeviceShardLoader(ds, 32, torch.device("cuda", 0), sampler=DistributedSampler(ds))
t.enable_resident_shard(shard)
st = t._state()
torch.cuda.synchronize()

def dev_time(fn, n=200):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)

print("device time of ONE replay (us, device idle before, includes launch latency):")
print("  dev-kind graph     ", round(dev_time(lambda: st['slots']['dev'].graph.replay()), 2))
print("  host-kind graph    ", round(dev_time(lambda: st['slots']['host'][0].graph.replay()), 2))
print("  resident graph     ", round(dev_time(lambda: st['resident']['graph'].replay()), 2))

def back_to_back(g, n=400):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n):
        g.replay()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n, (t1 - t0) * 1e6 / n

for name, g in (("dev", st['slots']['dev'].graph), ("host", st['slots']['host'][0].graph), ("resident", st['resident']['graph'])):
    d, h = back_to_back(g)
    print(f"back-to-back {name:9s}: device {d:6.2f} us/replay, host enqueue {h:6.2f} us/replay")

# host cost of each call in the pipelined host path
slots = st['slots']['host']
acc = {k: 0.0 for k in ("sync", "memmove", "replay", "record", "item", "total")}
N = 3000
prev = None
torch.cuda.synchronize()
T0 = time.perf_counter()
for i in range(N):
    slot = slots[i & 1]
    x, y = host[i % 64]
    t0 = time.perf_counter(); slot.done.synchronize()
    t1 = time.perf_counter(); ctypes.memmove(slot.src_ptr, x.data_ptr(), slot.src_bytes); ctypes.memmove(slot.tgt_ptr, y.data_ptr(), slot.tgt_bytes)
    t2 = time.perf_counter()
    if slot.stream is not None:
        with torch.cuda.stream(slot.stream):
            slot.graph.replay()
            t3 = time.perf_counter(); slot.done.record()
    else:
        slot.graph.replay()
        t3 = time.perf_counter(); slot.done.record()
    t4 = time.perf_counter()
    if prev is not None:
        prev.done.synchronize(); float(prev.host_loss_np[0])
    t5 = time.perf_counter()
    prev = slot
    acc["sync"] += t1 - t0; acc["memmove"] += t2 - t1; acc["replay"] += t3 - t2; acc["record"] += t4 - t3; acc["item"] += t5 - t4
torch.cuda.synchronize()
acc["total"] = time.perf_counter() - T0
print("pipelined host path, host us per step:", {k: round(v / N * 1e6, 2) for k, v in acc.items()})
# the same through Trainer._run_batch
T0 = time.perf_counter(); prev = None
for i in range(N):
    out = t._run_batch(*host[i % 64])
    if prev is not None:
        prev.item()
    prev = out
prev.item(); torch.cuda.synchronize()
print("Trainer._run_batch(host) + item of previous: us per step", round((time.perf_counter() - T0) / N * 1e6, 2))
T0 = time.perf_counter()
for i in range(N):
    t._run_resident_step()
torch.cuda.synchronize()
print("Trainer._run_resident_step: us per step", round((time.perf_counter() - T0) / N * 1e6, 2))
multigpu.destroy_process_group()
