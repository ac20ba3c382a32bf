"""Generates the golden fixtures in this directory.  Run ONCE in the build container (it needs
/root/reference, which does not exist on the GPU box); the outputs are committed.

    python tests/golden/make_golden.py

Fixtures
  bucket_assignment.json   inputs -> outputs of torch's own ``dist._compute_bucket_assignment_by_size``
                           (the dependency function DDP calls, torch:nn/parallel/distributed.py:1240)
  ref_trainer_ws2.pt       the REFERENCE's own ``multigpu.Trainer`` / ``load_train_objs`` /
                           ``prepare_dataloader`` / ``utils.MyTrainDataset`` (imported unmodified from
                           /root/reference) run for world_size=2 on gloo/CPU with only the names ``DDP``
                           and ``init_process_group`` rebound (no GPU here), seeds fixed, loss swapped to
                           MSE for the non-vacuous case and left as CE for the zero-grad case
  stock_ddp_mlp_ws3.pt     stock torch-DDP (gloo, ws=3) on the multi-tensor "mlp" workload: per-rank local
                           gradients and the gradients DDP produced, 2 steps
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------------
def bucket_cases():
    MiB = 1 << 20
    cases = []

    def add(name, numels, dtypes, limits, tensor_indices=None):
        cases.append(dict(name=name, numels=numels, dtypes=dtypes, limits=limits, tensor_indices=tensor_indices))

    add("survey_probe", [300000, 10, 70000, 200000, 5, 1000000, 8000000, 3], ["f32"] * 8, [MiB, 25 * MiB])
    add("reversed_small_cap", [3, 8000000, 1000000, 5, 200000, 70000, 10, 300000], ["f32"] * 8, [MiB, 4 * MiB])
    add("mixed_dtype", [10, 10, 300000, 5], ["f32", "bf16", "f32", "bf16"], [MiB, 4 * MiB])
    add("toy_linear", [20, 1], ["f32"] * 2, [MiB, 25 * MiB])
    add("single_limit_maxsize", [20, 1, 4096 * 4096, 4096], ["f32"] * 4, [sys.maxsize])
    add("wide4096", [4096 * 4096, 4096], ["f32"] * 2, [MiB, 25 * MiB])
    add("exact_boundary", [262144, 262144, 1, 6553600 - 1, 1, 7], ["f32"] * 6, [MiB, 25 * MiB])
    add("three_limits", [100] * 40, ["f32"] * 40, [400, 1200, 4000])
    order = [7, 6, 5, 4, 3, 2, 1, 0]
    add("ready_order", [300000, 10, 70000, 200000, 5, 1000000, 8000000, 3][::-1], ["f32"] * 8, [MiB, 25 * MiB], order)
    order2 = [3, 1, 2, 0]
    add("ready_order_mixed", [5, 10, 300000, 10], ["bf16", "bf16", "f32", "f32"], [MiB, 4 * MiB], order2)
    # real architectures (shapes only, meta device)
    from torchvision.models import resnet50

    with torch.device("meta"):
        rn = resnet50()
    rn_numels = [p.numel() for p in rn.parameters()]
    add("resnet50_definition_order", rn_numels, ["f32"] * len(rn_numels), [MiB, 25 * MiB])
    rev = list(range(len(rn_numels)))[::-1]
    add("resnet50_reverse_ready_order", [rn_numels[i] for i in rev], ["f32"] * len(rn_numels), [MiB, 25 * MiB], rev)
    import b200ddp  # noqa: F401
    from b200ddp.workloads import GPT2Shaped

    with torch.device("meta"):
        g = GPT2Shaped()
    g_numels = [p.numel() for p in g.parameters()]
    rev = list(range(len(g_numels)))[::-1]
    add("gpt2m_bf16_reverse_ready_order", [g_numels[i] for i in rev], ["bf16"] * len(g_numels), [MiB, 25 * MiB], rev)

    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
    for c in cases:
        with torch.device("meta"):
            tensors = [torch.empty(n, dtype=tdt[d]) for n, d in zip(c["numels"], c["dtypes"])]
        kw = {}
        if c["tensor_indices"] is not None:
            kw["tensor_indices"] = c["tensor_indices"]
        buckets, limits = dist._compute_bucket_assignment_by_size(tensors, c["limits"], [False] * len(tensors), **kw)
        c["expect_buckets"] = [list(b) for b in buckets]
        c["expect_limits"] = list(limits)
    with open(os.path.join(HERE, "bucket_assignment.json"), "w") as f:
        json.dump(dict(torch=str(torch.__version__), cases=cases), f)
    print("bucket_assignment.json:", len(cases), "cases")


# ------------------------------------------------------------------------------------------------------
def _ref_trainer_worker(rank, world, port, loss_name, steps, ret):
    sys.path.insert(0, REF)
    import multigpu as ref  # the reference's own script, unmodified

    # the only rebinding: no GPU in this container -> stock DDP without device_ids, gloo instead of nccl
    ref.DDP = lambda model, device_ids=None: torch.nn.parallel.DistributedDataParallel(model)
    ref.init_process_group = lambda backend, rank, world_size: dist.init_process_group(
        "gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    if loss_name == "mse":  # the scripts' CE on one logit has identically zero gradients
        ref.torch.nn.CrossEntropyLoss = torch.nn.MSELoss
    ref.ddp_setup(rank, world)
    torch.manual_seed(1234 + rank)  # the reference seeds nothing (Q2): per-rank init + data, like it
    dataset, model, optimizer = ref.load_train_objs()
    loader = ref.prepare_dataloader(dataset, batch_size=32)
    trainer = ref.Trainer(model, loader, optimizer, "cpu", save_every=10**9)
    rec = dict(x=[], y=[], w_before=[], ddp_grads=[], w_after=[], local_grads=[])
    inner = trainer._run_batch
    loss_fn = torch.nn.MSELoss() if loss_name == "mse" else torch.nn.functional.cross_entropy

    def recording_run_batch(source, targets):
        if len(rec["x"]) < steps:
            rec["x"].append(source.clone())
            rec["y"].append(targets.clone())
            rec["w_before"].append([p.detach().clone() for p in model.parameters()])
            shadow = torch.nn.Linear(20, 1)
            shadow.load_state_dict(model.state_dict())
            loss_fn(shadow(source), targets).backward()
            rec["local_grads"].append([p.grad.clone() for p in shadow.parameters()])
        inner(source, targets)
        if len(rec["ddp_grads"]) < steps:
            rec["ddp_grads"].append([p.grad.detach().clone() for p in model.parameters()])
            rec["w_after"].append([p.detach().clone() for p in model.parameters()])

    trainer._run_batch = recording_run_batch
    trainer.train(1)
    rec["sampler_indices"] = list(iter(loader.sampler))[:64]
    ret[rank] = rec
    dist.destroy_process_group()


def ref_trainer():
    from oracle.ref_ddp import free_port

    out = {}
    for loss_name, steps in (("mse", 4), ("ce", 1)):
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_ref_trainer_worker, args=(2, free_port(), loss_name, steps, ret), nprocs=2, join=True)
            out[loss_name] = {r: dict(ret[r]) for r in range(2)}
    torch.save(dict(torch=str(torch.__version__), world=2, seed_rule="torch.manual_seed(1234 + rank) before load_train_objs",
                    cases=out), os.path.join(HERE, "ref_trainer_ws2.pt"))
    g = out["mse"][0]["ddp_grads"][0]
    print("ref_trainer_ws2.pt: mse grad absmax", max(float(t.abs().max()) for t in g),
          "| ce grad absmax", max(float(t.abs().max()) for t in out["ce"][0]["ddp_grads"][0]))


def stock_mlp():
    from oracle.ref_ddp import run_stock_ddp

    out = run_stock_ddp(3, "mlp", steps=2, batch_size=16, seed=77)
    # DDP leaves identical gradients on every rank: keep rank 0's, plus every rank's local gradients
    same = all(torch.equal(a, b) for r in (1, 2) for s in range(2)
               for a, b in zip(out[0]["record"]["ddp_grads"][s], out[r]["record"]["ddp_grads"][s]))
    assert same, "stock DDP produced different gradients on different ranks"
    keep = {r: dict(local_grads=out[r]["record"]["local_grads"]) for r in out}
    keep[0]["ddp_grads"] = out[0]["record"]["ddp_grads"]
    torch.save(dict(torch=str(torch.__version__), world=3, workload="mlp", seed=77, batch_size=16, ranks=keep),
               os.path.join(HERE, "stock_ddp_mlp_ws3.pt"))
    print("stock_ddp_mlp_ws3.pt written")


if __name__ == "__main__":
    only = sys.argv[1:] or ["buckets", "trainer", "mlp"]
    if "buckets" in only:
        bucket_cases()
    if "trainer" in only:
        ref_trainer()
    if "mlp" in only:
        stock_mlp()
