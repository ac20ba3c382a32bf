"""Fresh-process probes of CUDA-graph capture of the toy step (which hook loop, which stream the DDP was built on)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, copy
sys.path.insert(0, %r)
import torch
import b200ddp
from b200ddp import multigpu
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29655")
multigpu.ddp_setup(0, 1)
multigpu.Trainer.loss_fn = staticmethod(torch.nn.MSELoss())
torch.manual_seed(0)
model = torch.nn.Linear(20, 1)
opt = torch.optim.SGD(model.parameters(), lr=1e-2)
t = multigpu.Trainer(model, [], opt, 0, save_every=10**9)
x, y = torch.rand(32, 20).cuda(), torch.rand(32, 1).cuda()
hx, hy = torch.rand(32, 20).pin_memory(), torch.rand(32, 1).pin_memory()
kind = os.environ.get("KIND", "dev")
for i in range(8):
    out = t._run_batch(x, y) if kind == "dev" else t._run_batch(hx, hy)
torch.cuda.synchronize()
st = t._state()
print("RESULT native=%%s kind=%%s slots=%%s failed=%%s" %% (t.model.reducer._native is not None, kind, list(st["slots"]), st["failed"]))
multigpu.destroy_process_group()
''' % ROOT
for native in ("1", "0"):
    for kind in ("dev", "host"):
        env = dict(os.environ, B200DDP_NATIVE=native, KIND=kind, B200DDP_TIMEOUT_MS="20000")
        env.pop("B200DDP_GRAPH", None)
        p = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=200)
        lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
        print(f"NATIVE={native} KIND={kind} rc={p.returncode}:", lines[-1] if lines else p.stderr[-600:], flush=True)
