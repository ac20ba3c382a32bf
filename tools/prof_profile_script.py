"""Drive ONE capture through the reference-facing profiling surface: the `multigpu_profile.Trainer(
..., profile=True)` of ref:multigpu_profile.py:30-91, whose CudaProfilerWindow opens the CUDA profiler range for the
reference's wait=1 / warmup=1 / active=5 schedule.  ONE GPU, world 1 (ncu replays kernels in isolation):

    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:multi_tensor_copy \\
        -c 20 -o gpurun_out/r2_profile_script python tools/prof_profile_script.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import b200ddp  # noqa: E402,F401
from b200ddp import destroy_process_group, multigpu_profile as mp  # noqa: E402
from b200ddp.utils import MyRandomDataset  # noqa: E402

os.environ.setdefault("MASTER_PORT", "12377")
mp.ddp_setup(0, 1)
_, net, sgd = mp.load_train_objs()
train_set = MyRandomDataset(32 * 8, (3, 224, 224), seed=0)  # 8 steps: 1 wait + 1 warm-up + 5 active + 1
trainer = mp.Trainer(net, train_set, sgd, device=0, batch_size=32, profile=True)
trainer.run_epoch(1)
torch.cuda.synchronize()
w = trainer.profiler
print(f"profile window covered steps [{w.first}, {w.last}) of {w.n}; b200ddp launches: "
      f"{trainer.model.engine.comm.launch_count}")
destroy_process_group()
