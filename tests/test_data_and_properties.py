"""CPU tests: the device-resident shard loader serves exactly the DataLoader's batches; bucketing agrees with
the oracle on random inputs (property test); world-size-2 gloo checks of buffer broadcast and no_sync."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings
from hypothesis import strategies as st

from b200ddp.bucketing import compute_bucket_assignment_by_size
from oracle import bucket_oracle as bo


@settings(max_examples=200, deadline=None)
@given(
    sizes=st.lists(st.integers(min_value=0, max_value=3_000_000), min_size=0, max_size=40),
    limits=st.lists(st.integers(min_value=1, max_value=4_000_000), min_size=1, max_size=3),
    keyseed=st.integers(min_value=0, max_value=3),
    ready=st.booleans(),
    data=st.data(),
)
def test_bucketing_matches_oracle_on_random_inputs(sizes, limits, keyseed, ready, data):
    keys = [("f32", "d0") if (i * 7 + keyseed) % 3 or keyseed == 0 else ("bf16", "d0") for i in range(len(sizes))]
    order = data.draw(st.permutations(range(len(sizes)))) if ready else None
    got = compute_bucket_assignment_by_size(sizes, keys, limits, order)
    want = bo.assign_buckets(sizes, keys, limits, order)
    assert got == want
    flat = sorted(i for b in got[0] for i in b)
    assert flat == list(range(len(sizes)))  # a partition of the tensors
    assert len(got[0]) == len(got[1]) and all(l in limits for l in got[1])


def _shard_worker(rank, world, port, ret):
    from b200ddp import multigpu, utils

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ds = utils.MyTrainDataset(2048 + 5, seed=3)  # not divisible: the sampler pads
        loader = multigpu.prepare_dataloader(ds, 32)
        fast = utils.DeviceShardLoader(ds, 32, "cpu")
        same = len(loader) == len(fast)
        for (a, b), (c, d) in zip(loader, fast):
            same = same and torch.equal(a, c) and torch.equal(b, d)
        ret[rank] = dict(same=same, n=len(fast), rows=int(fast.x.shape[0]))
    finally:
        dist.destroy_process_group()


def _bn_nosync_worker(rank, world, port, ret):
    """BatchNorm buffers: rank 0's running stats are broadcast before every synchronised forward
    (torch:nn/parallel/distributed.py:1590-1591); no_sync(): gradients accumulate locally and the next
    synchronised backward reduces the accumulated sum — both compared with stock DDP on the same inputs."""
    import copy

    from b200ddp import DDP
    from tests.helpers import GlooEngine, rel_err

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        torch.manual_seed(5 + rank)  # different init per rank: construction-time broadcast must fix it
        base = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
        ours_m, ref_m = copy.deepcopy(base), copy.deepcopy(base)
        ours = DDP(ours_m, _engine=GlooEngine())
        ref = torch.nn.parallel.DistributedDataParallel(ref_m)
        gen = torch.Generator().manual_seed(100 + rank)
        worst = 0.0
        buffers_equal = True
        for step in range(4):
            x, y = torch.rand(16, 6, generator=gen), torch.rand(16, 3, generator=gen)
            ours_m.zero_grad()
            ref_m.zero_grad()
            if step % 2 == 0:  # an un-synchronised micro-batch first
                x2, y2 = torch.rand(16, 6, generator=gen), torch.rand(16, 3, generator=gen)
                with ours.no_sync():
                    torch.nn.functional.mse_loss(ours(x2), y2).backward()
                with ref.no_sync():
                    torch.nn.functional.mse_loss(ref(x2), y2).backward()
            torch.nn.functional.mse_loss(ours(x), y).backward()
            torch.nn.functional.mse_loss(ref(x), y).backward()
            for p, q in zip(ours_m.parameters(), ref_m.parameters()):
                worst = max(worst, rel_err(p.grad, q.grad))
            for a, b in zip(ours_m.buffers(), ref_m.buffers()):
                buffers_equal = buffers_equal and torch.allclose(a.float(), b.float(), rtol=0, atol=1e-7)
        ret[rank] = dict(worst=worst, buffers_equal=buffers_equal)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world):
    from oracle.ref_ddp import free_port

    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(fn, args=(world, free_port(), ret), nprocs=world, join=True)
        return {r: dict(ret[r]) for r in range(world)}


@pytest.mark.timeout(300)
def test_device_shard_loader_serves_the_dataloaders_batches():
    out = _spawn(_shard_worker, 2)
    for r in range(2):
        assert out[r]["same"], f"rank {r}: device-resident batches differ from prepare_dataloader's"
        assert out[r]["rows"] == 1027 and out[r]["n"] == 33  # ceil(2053 / 2) rows, last batch of 3 kept


@pytest.mark.timeout(300)
def test_ws2_batchnorm_buffers_and_no_sync_match_stock_ddp():
    out = _spawn(_bn_nosync_worker, 2)
    for r in range(2):
        assert out[r]["worst"] <= 1e-6, out[r]
        assert out[r]["buffers_equal"]


@settings(max_examples=200, deadline=None)
@given(
    numels=st.lists(st.integers(min_value=0, max_value=100_000), min_size=1, max_size=30),
    esize=st.sampled_from([2, 4]),
    start=st.integers(min_value=0, max_value=1 << 20),
    modes=st.lists(st.sampled_from(["plain", "oop", "region"]), min_size=1, max_size=6),
)
def test_arena_layout_invariants(numels, esize, start, modes):
    """Every tensor starts on a 16-byte boundary inside its bucket without overlapping its neighbour; bucket
    regions are 4 KiB aligned, never overlap and never precede the staging region."""
    from b200ddp.bucketing import BucketLayout, layout_bucket

    spec = layout_bucket(0, list(range(len(numels))), numels, dtype_code=0 if esize == 4 else 1, elem_size=esize)
    epv = 16 // esize
    end = 0
    for off, n in zip(spec.offsets_el, spec.lengths_el):
        assert off % epv == 0 and off >= end
        end = off + n
    assert spec.nbytes % 16 == 0 and spec.nbytes >= max(16, end * esize) and spec.nbytes < end * esize + 16 * len(numels) + 16
    lay = BucketLayout(start=start)
    assert lay.total_bytes >= start and lay.total_bytes % 4096 == 0
    spans = []
    for k, mode in enumerate(modes):
        b = layout_bucket(k, [0], [numels[0] + k], dtype_code=0, elem_size=4)
        before = lay.total_bytes
        if mode == "region":
            region = 4 * 8 * b.nbytes
            lay.place(b, region_bytes=region)
            need = region
        else:
            lay.place(b, out_of_place=(mode == "oop"))
            need = b.nbytes * (2 if mode == "oop" else 1)
        assert b.arena_off == before and b.arena_off % 4096 == 0
        assert lay.total_bytes - before >= need and lay.total_bytes % 4096 == 0
        if mode == "oop":
            assert b.out_off >= b.arena_off + b.nbytes and b.out_off % 4096 == 0
        else:
            assert b.out_off == b.arena_off
        spans.append((before, lay.total_bytes))
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
