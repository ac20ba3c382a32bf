"""One-process-per-GPU workers for the multi-GPU tests (spawned by tests/test_gpu_ddp.py, also runnable
under torchrun: ``torchrun --nproc-per-node N tests/gpu_mp_worker.py <fn> [k=v ...]``)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def raw_allreduce(rank, world, mem="vmm"):
    """Every algorithm on every dtype at ragged sizes: bit-exact vs the oracle for the P2P algorithms
    (fixed summation order), <= 1e-6 rel for NVLS (the switch chooses the order), and vs NCCL."""
    from b200ddp import _lib
    from b200ddp.comm import P2PComm
    from oracle import allreduce_oracle as ao

    os.environ["B200DDP_MEM"] = mem
    comm = P2PComm.from_process_group(64 << 20, rank)
    comm.set_timeout_ms(20000)
    res = dict(ok=True, mem_mode=comm.mem_mode, has_multicast=comm.has_multicast, fails=[])
    algos = [("oneshot", 1), ("twoshot", 2), ("tma", 3)] + ([("nvls", 4)] if comm.has_multicast else [])
    dts = {"f32": (torch.float32, _lib.F32), "bf16": (torch.bfloat16, _lib.BF16)}
    stream = torch.cuda.Stream()
    for dname, (tdt, code) in dts.items():
        es = 2 if dname != "f32" else 4
        for n_el in (16 // es, (16 // es) * 1001, (4 << 20) // es + 16 // es * 3):
            nbytes = n_el * es
            gens = [torch.Generator().manual_seed(1000 * n_el % 7919 + q) for q in range(world)]
            all_bufs = [(torch.rand(n_el, generator=g) * 2 - 1).to(tdt) for g in gens]
            mine = all_bufs[rank].cuda()

            def store(t):
                t = t.cpu().contiguous()
                return t.numpy().ravel() if dname == "f32" else t.view(torch.int16).numpy().view(np.uint16).ravel()

            want = ao.allreduce_sum([store(b) for b in all_bufs], dname)
            nccl = mine.clone()
            dist.all_reduce(nccl)
            for aname, algo in algos:
                in_off, out_off = 0, 32 << 20
                comm.view(in_off, nbytes, tdt).copy_(mine)
                torch.cuda.synchronize()
                dist.barrier()
                comm.allreduce(in_off, out_off if aname == "oneshot" else in_off, nbytes, code, algo, 0,
                               stream.cuda_stream)
                stream.synchronize()
                comm.poll_error()
                got_t = comm.view(out_off if aname == "oneshot" else in_off, nbytes, tdt).clone()
                got = store(got_t)
                if aname == "nvls":
                    a = ao.to_f32(got, dname).astype(np.float64)
                    b = ao.to_f32(want, dname).astype(np.float64)
                    tol = 1e-6 if dname == "f32" else 2 ** -7
                    ok = np.abs(a - b).max() / np.abs(b).max() <= tol
                else:
                    ok = np.array_equal(got, want)
                a = got_t.float().double()
                b = nccl.float().double()
                rel_nccl = float((a - b).abs().max() / b.abs().max())
                ok = ok and rel_nccl <= (1e-6 if dname == "f32" else 2 ** -6)
                if not ok:
                    res["ok"] = False
                    res["fails"].append(f"{aname}/{dname}/{n_el}: rel_vs_nccl={rel_nccl:.3e}")
                dist.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    comm.destroy()
    return res


def ddp_vs_nccl(rank, world, workload="mlp", steps=4, bucket_cap_mb=None, views=False, batch=16, regrow_at=None):
    """Our DDP and stock torch-DDP/NCCL side by side: same seeded weights, same per-rank batches."""
    import copy

    from b200ddp import DDP, workloads

    if workload == "resnet50":  # parity needs the two replicas' backward kernels to be the same ones
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    wl = workloads.get(workload)
    dev = torch.device("cuda", rank)
    torch.manual_seed(4321)
    base = wl.build().to(dev)
    ours_m, ref_m = copy.deepcopy(base), copy.deepcopy(base)
    kw = {} if bucket_cap_mb is None else {"bucket_cap_mb": bucket_cap_mb}
    ours = DDP(ours_m, device_ids=[rank], gradient_as_bucket_view=bool(views), **kw)
    ref = torch.nn.parallel.DistributedDataParallel(ref_m, device_ids=[rank], gradient_as_bucket_view=bool(views), **kw)
    o1 = torch.optim.SGD(ours_m.parameters(), lr=1e-3)
    o2 = torch.optim.SGD(ref_m.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(99 + rank)
    worst = 0.0
    for step in range(steps):
        x, y = wl.batch(batch, dev, gen)
        if regrow_at is not None and step == regrow_at:
            # what a re-bucketing that outgrows the arena does (P2PEngine._grow_arena, collective): a new, larger
            # communicator; the Reducer re-installs its buckets on it and re-binds the C++ hook loop
            o1.zero_grad()
            ours.engine._grow_arena(2 * ours.engine.comm.arena_bytes)
            ours.reducer._install([list(b.param_indices) for b in ours.reducer.buckets])
        ref_m.load_state_dict(ours_m.state_dict())  # same weights going into every step
        o1.zero_grad()
        o2.zero_grad()
        wl.loss(ours(x), y).backward()
        wl.loss(ref(x), y).backward()
        torch.cuda.synchronize()
        for p, q in zip(ours_m.parameters(), ref_m.parameters()):
            denom = float(q.grad.double().abs().max())
            num = float((p.grad.double() - q.grad.double()).abs().max())
            worst = max(worst, num / denom if denom > 0 else num)
        o1.step()
    flat = torch.cat([p.detach().reshape(-1) for p in ours_m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    res = dict(worst_rel=worst, weights_equal_across_ranks=same, buckets=ours.reducer.bucket_sizes(),
               fused=[b.fused for b in ours.reducer.buckets], algos=[b.algo for b in ours.reducer.buckets],
               mem_mode=ours.engine.comm.mem_mode, has_multicast=ours.engine.comm.has_multicast)
    return res


def hook_vs_nccl(rank, world, workload="slices", steps=3):
    """Stock torch DDP with our allreduce as its comm hook (hooks.allreduce_hook) against stock torch DDP: the
    dependency's own plugin API for the step (SURVEY §8 b).  Same weights, same batches."""
    import copy

    from b200ddp import hooks, workloads

    wl = workloads.get(workload)
    dev = torch.device("cuda", rank)
    torch.manual_seed(4321)
    base = wl.build().to(dev)
    a, b = copy.deepcopy(base), copy.deepcopy(base)
    stock = torch.nn.parallel.DistributedDataParallel(a, device_ids=[rank])
    hooked = torch.nn.parallel.DistributedDataParallel(b, device_ids=[rank])
    state = hooks.AllreduceState(rank, staging_mb=16)  # the 64 MiB first-iteration bucket is chunked
    hooked.register_comm_hook(state, hooks.allreduce_hook)
    gen = torch.Generator().manual_seed(99 + rank)
    worst = 0.0
    for step in range(steps):
        x, y = wl.batch(16, dev, gen)
        for m in (stock, hooked):
            m.zero_grad()
            wl.loss(m(x), y).backward()
        torch.cuda.synchronize()
        for p, q in zip(b.parameters(), a.parameters()):
            denom = float(q.grad.double().abs().max())
            num = float((p.grad.double() - q.grad.double()).abs().max())
            worst = max(worst, num / denom if denom > 0 else num)
    res = dict(worst_rel=worst, buckets_reduced=state.buckets_reduced, has_multicast=state.comm.has_multicast)
    torch.cuda.synchronize()
    dist.barrier()
    del stock, hooked
    state.close()
    return res


def fused_sgd_vs_nccl(rank, world, workload="mlp", steps=5, bucket_cap_mb=None):
    """Our DDP with the optimizer epilogue (fuse_optimizer) against stock torch-DDP/NCCL + torch.optim.SGD: the
    weights after `steps` steps (small fused buckets and, for `slices`, 25 MiB buckets through the unflatten kernel)."""
    import copy

    from b200ddp import DDP, workloads

    wl = workloads.get(workload)
    dev = torch.device("cuda", rank)
    torch.manual_seed(4321)
    base = wl.build().to(dev)
    ours_m, ref_m = copy.deepcopy(base), copy.deepcopy(base)
    kw = {} if bucket_cap_mb is None else {"bucket_cap_mb": bucket_cap_mb}
    ours = DDP(ours_m, device_ids=[rank], **kw)
    ref = torch.nn.parallel.DistributedDataParallel(ref_m, device_ids=[rank], **kw)
    o1 = torch.optim.SGD(ours_m.parameters(), lr=1e-2)
    o2 = torch.optim.SGD(ref_m.parameters(), lr=1e-2)
    fused = ours.fuse_optimizer(o1)
    gen = torch.Generator().manual_seed(99 + rank)
    for step in range(steps):
        x, y = wl.batch(16, dev, gen)
        for m, o in ((ours, o1), (ref, o2)):
            o.zero_grad()
            wl.loss(m(x), y).backward()
            o.step()
    torch.cuda.synchronize()
    worst = 0.0
    for p, q in zip(ours_m.parameters(), ref_m.parameters()):
        denom = float(q.double().abs().max())
        num = float((p.double() - q.double()).abs().max())
        worst = max(worst, num / denom if denom > 0 else num)
    flat = torch.cat([p.detach().reshape(-1) for p in ours_m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    return dict(fused=bool(fused), worst_rel_weights=worst, weights_equal_across_ranks=all(torch.equal(g, gathered[0]) for g in gathered),
                fused_buckets=[b.fused for b in ours.reducer.buckets])


FUNCS = dict(raw_allreduce=raw_allreduce, ddp_vs_nccl=ddp_vs_nccl, hook_vs_nccl=hook_vs_nccl,
             fused_sgd_vs_nccl=fused_sgd_vs_nccl)


def entry(rank, world, port, fn_name, kw, ret):
    from b200ddp import destroy_process_group, init_process_group

    os.environ["LOCAL_RANK"] = str(rank)
    os.environ.setdefault("B200DDP_TIMEOUT_MS", "30000")
    torch.cuda.set_device(rank)
    init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ret[rank] = FUNCS[fn_name](rank, world, **kw)
    finally:
        destroy_process_group()


if __name__ == "__main__":
    from b200ddp import destroy_process_group, init_process_group

    fn = sys.argv[1]
    kw = dict(a.split("=", 1) for a in sys.argv[2:])
    for k, v in list(kw.items()):
        try:
            kw[k] = json.loads(v)
        except ValueError:
            pass
    init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    try:
        out = FUNCS[fn](rank, world, **kw)
        print(f"[rank {rank}] {json.dumps(out)}", flush=True)
    finally:
        destroy_process_group()
