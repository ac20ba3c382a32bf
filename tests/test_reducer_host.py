"""Host logic of the custom Reducer and the DDP wrapper, exercised on CPU with test engines
(tests/helpers.py): hook/pending/ordering protocol, one-time rebuild, error paths, surface."""
import inspect

import pytest
import torch
import torch.nn as nn

from b200ddp import DDP
from b200ddp.bucketing import BucketLayout, layout_bucket
from tests.helpers import RecordingEngine


def _mlp():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(20, 96), nn.ReLU(), nn.Linear(96, 96), nn.ReLU(), nn.Linear(96, 7))


def _step(ddp, x=None):
    x = torch.randn(8, 20) if x is None else x
    ddp.zero_grad()
    ddp(x).square().mean().backward()


def test_initial_buckets_follow_reverse_definition_order_and_rebuild_once():
    eng = RecordingEngine()
    ddp = DDP(_mlp(), bucket_cap_mb=0.01, _engine=eng)  # ~10 KB cap -> several buckets
    sizes0 = [b.param_indices for b in eng.plans[0]]
    # construction: greedy over definition order, list reversed (bucket 0 = last parameters)
    flat = [i for b in sizes0 for i in b]
    assert sorted(flat) == list(range(6))
    assert sizes0[0][-1] == 5 and sizes0[-1][0] == 0
    _step(ddp)
    # launches happened strictly in bucket-index order and covered every bucket
    assert eng.launched == list(range(len(sizes0)))
    assert eng.finalized == 1
    # one rebuild after iteration 0, in gradient-ready order (reverse of definition for a chain); like torch it
    # is carried out by the NEXT forward (distributed.py:1584), so that gradients served as bucket views stay
    # valid for the optimizer step in between
    assert len(eng.plans) == 1 and not ddp.reducer.has_rebuilt
    eng.launched.clear()
    _step(ddp)
    assert len(eng.plans) == 2 and ddp.reducer.has_rebuilt
    ready = [i for b in eng.plans[1] for i in b.param_indices]
    assert ready == [5, 4, 3, 2, 1, 0]
    assert eng.launched == list(range(len(eng.plans[1])))
    eng.launched.clear()
    _step(ddp)
    assert eng.launched == list(range(len(eng.plans[1])))
    assert len(eng.plans) == 2  # never rebuilt again


def test_default_caps_give_one_bucket_for_toy_model():
    eng = RecordingEngine()
    ddp = DDP(nn.Linear(20, 1), _engine=eng)
    assert [b.param_indices for b in eng.plans[0]] == [[0, 1]]
    assert ddp.reducer.bucket_sizes() == [84]  # toy = 1 bucket x 84 B


def test_grads_land_in_original_parameters_and_surface():
    model = _mlp()
    opt_params = list(model.parameters())
    ddp = DDP(model, device_ids=None, _engine=RecordingEngine())
    assert ddp.module is model
    assert all(a is b for a, b in zip(opt_params, ddp.module.parameters()))
    assert all(k.startswith("module.") for k in ddp.state_dict())  # ref:multigpu_profile.py:77
    assert list(ddp.module.state_dict()) == [k[len("module."):] for k in ddp.state_dict()]
    _step(ddp)
    assert all(p.grad is not None for p in opt_params)
    ref = _mlp()
    x = torch.randn(8, 20)
    _step(ddp, x)
    ref(x).square().mean().backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.equal(p.grad, q.grad)  # ws=1 recording engine leaves grads untouched


def test_out_of_order_ready_still_launches_in_index_order():
    eng = RecordingEngine()
    model = _mlp()
    ddp = DDP(model, bucket_cap_mb=0.01, _engine=eng)
    _step(ddp)  # rebuild happens here
    eng.launched.clear()
    # fire hooks by hand in an order where a later bucket completes first
    ddp(torch.randn(2, 20))
    red = ddp.reducer
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    last = red.buckets[-1]
    for i in last.param_indices:
        red._param_seen[i] = True
        red._pending[last.index] -= 1
    red._bucket_ready[last.index] = red._pending[last.index] == 0
    assert eng.launched == []  # bucket 0 not ready yet -> nothing may launch (strict order, reducer.hpp:231)


def test_unused_parameter_raises_like_ddp():
    class Partial(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(4, 4), nn.Linear(4, 4)

        def forward(self, x):
            return self.a(x)

    ddp = DDP(Partial(), _engine=RecordingEngine())
    with pytest.raises(RuntimeError, match="never produced a gradient"):
        ddp(torch.randn(2, 4)).sum().backward()
    with pytest.raises(RuntimeError, match="finished reduction in the prior iteration"):
        ddp(torch.randn(2, 4))


def test_no_sync_skips_reduction_and_accumulates():
    eng = RecordingEngine()
    ddp = DDP(_mlp(), _engine=eng)
    x = torch.randn(4, 20)
    with ddp.no_sync():
        ddp(x).sum().backward()
    assert eng.launched == [] and eng.finalized == 0
    ddp(x).sum().backward()
    assert eng.launched and eng.finalized == 1


def test_backward_without_forward_does_not_reduce():
    eng = RecordingEngine()
    model = _mlp()
    ddp = DDP(model, _engine=eng)
    model(torch.randn(2, 20)).sum().backward()  # inner module called directly: reducer not armed
    assert eng.launched == []


def test_constructor_argument_errors():
    with pytest.raises(NotImplementedError):
        DDP(_mlp(), find_unused_parameters=True, _engine=RecordingEngine())
    with pytest.raises(ValueError):
        DDP(_mlp(), device_ids=[0, 1], _engine=RecordingEngine())
    frozen = nn.Linear(2, 2)
    for p in frozen.parameters():
        p.requires_grad_(False)
    with pytest.raises(RuntimeError, match="doesn't have any parameter"):
        DDP(frozen, _engine=RecordingEngine())


def test_bucket_layout_alignment():
    spec = layout_bucket(0, [0, 1, 2], [20, 1, 5], dtype_code=0, elem_size=4)
    assert spec.offsets_el == [0, 20, 24] and spec.nbytes == 32 * 4
    spec16 = layout_bucket(0, [0, 1], [3, 9], dtype_code=1, elem_size=2)
    assert spec16.offsets_el == [0, 8] and spec16.nbytes == 48
    assert BucketLayout(start=5000).total_bytes == 8192
    lay = BucketLayout()
    lay.place(spec, out_of_place=True)
    assert spec.arena_off == 0 and spec.out_off == 4096 and lay.total_bytes == 8192
    lay.place(spec16, out_of_place=False)
    assert spec16.arena_off == spec16.out_off == 8192


def test_script_surface_matches_reference_signatures():
    """Names and signatures of the reference scripts (ref:multigpu.py, ref:multigpu_torchrun.py, ref:utils.py)."""
    from b200ddp import multigpu, multigpu_torchrun, utils

    def params(fn):
        return list(inspect.signature(fn).parameters)

    assert params(multigpu.ddp_setup) == ["rank", "world_size"]
    assert params(multigpu.Trainer.__init__) == ["self", "model", "train_data", "optimizer", "gpu_id", "save_every"]
    assert params(multigpu.main) == ["rank", "world_size", "save_every", "total_epochs"]
    assert params(multigpu.prepare_dataloader) == ["dataset", "batch_size"]
    assert params(multigpu.load_train_objs) == []
    assert params(multigpu_torchrun.ddp_setup) == []
    assert params(multigpu_torchrun.Trainer.__init__) == ["self", "model", "train_data", "optimizer", "save_every",
                                                          "snapshot_path"]
    assert params(multigpu_torchrun.main) == ["save_every", "total_epochs", "snapshot_path"]
    for cls in (multigpu.Trainer, multigpu_torchrun.Trainer):
        for m in ("_run_batch", "_run_epoch", "train"):
            assert callable(getattr(cls, m))
    assert hasattr(multigpu.Trainer, "_save_checkpoint") and hasattr(multigpu_torchrun.Trainer, "_save_snapshot")
    ds, model, opt = multigpu.load_train_objs()
    assert len(ds) == 2048 and isinstance(model, nn.Linear) and isinstance(opt, torch.optim.SGD)
    x, y = ds[0]
    assert x.shape == (20,) and y.shape == (1,) and 0 <= float(x.min()) and float(x.max()) < 1
    rd = utils.MyRandomDataset(4, (3, 8, 8))
    a, b = rd[0]
    assert a.shape == (3, 8, 8) and b.shape == (1000,)
    loader = multigpu.prepare_dataloader(utils.MyTrainDataset(64, seed=1), 32) if torch.distributed.is_initialized() else None
    assert loader is None or loader.batch_size == 32


def test_mixed_dtypes_frozen_parameters_and_env_cap(monkeypatch):
    """Buckets never mix dtypes (BucketKey = dtype+device), frozen parameters are ignored, and
    B200DDP_BUCKET_MB overrides the default cap like bucket_cap_mb does."""
    class Mixed(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Linear(16, 16)
            self.b = nn.Linear(16, 16).to(torch.bfloat16)
            self.c = nn.Linear(16, 4)
            self.frozen = nn.Linear(4, 4)
            for p in self.frozen.parameters():
                p.requires_grad_(False)

        def forward(self, x):
            return self.frozen(self.c(self.b(self.a(x).to(torch.bfloat16)).float()))

    eng = RecordingEngine()
    ddp = DDP(Mixed(), _engine=eng)
    params = ddp.reducer.params
    assert len(params) == 6  # the frozen layer's two tensors are not reduced
    for spec in eng.plans[0]:
        dts = {params[i].dtype for i in spec.param_indices}
        assert len(dts) == 1
        assert spec.elem_size == params[spec.param_indices[0]].element_size()
    assert {spec.dtype_code for spec in eng.plans[0]} == {0, 1}
    ddp(torch.randn(3, 16)).sum().backward()
    assert sorted(eng.launched) == list(range(len(eng.plans[0])))
    assert all(p.grad is None for p in ddp.module.frozen.parameters())

    monkeypatch.setenv("B200DDP_BUCKET_MB", "0.001")
    small = DDP(_mlp(), _engine=RecordingEngine())
    assert small.bucket_bytes_cap == int(0.001 * 1024 * 1024)
    from oracle import bucket_oracle as bo

    sizes = [p.numel() * 4 for p in small.reducer.params]
    cap = small.bucket_bytes_cap
    want = bo.ddp_initial_buckets(sizes, ["k"] * len(sizes), first_cap=cap, cap=cap)
    assert [b.param_indices for b in small.reducer.buckets] == want and len(want) == 4


def test_buffer_sync_flag_follows_the_previous_forward():
    """require_forward_param_sync mirrors torch (distributed.py:1637-1650): buffers are broadcast before a
    forward iff the previous forward was a gradient-synchronising one."""
    class Counting(RecordingEngine):
        def __init__(self):
            super().__init__(world_size=2)
            self.bcasts = 0

        def broadcast_tensors(self, tensors, root=0):
            self.bcasts += 1

    eng = Counting()
    model = nn.Sequential(nn.Linear(4, 4), nn.BatchNorm1d(4))
    ddp = DDP(model, _engine=eng)
    assert eng.bcasts == 1  # construction-time parameter + buffer sync
    x = torch.randn(8, 4)
    ddp(x).sum().backward()
    assert eng.bcasts == 2  # first forward syncs buffers
    with ddp.no_sync():
        ddp(x).sum().backward()
    assert eng.bcasts == 3  # previous forward was synchronising
    ddp(x).sum().backward()
    assert eng.bcasts == 3  # previous forward was inside no_sync(): no buffer sync now
    with torch.no_grad():
        ddp(x)
    assert eng.bcasts == 4
    ddp(x).sum().backward()
    assert eng.bcasts == 4  # previous forward ran without grad
    plain = DDP(nn.Sequential(nn.Linear(4, 4), nn.BatchNorm1d(4)), broadcast_buffers=False, _engine=Counting())
    plain(x).sum().backward()
    assert plain.engine.bcasts == 1


def test_arena_is_sized_from_the_construction_time_plan():
    """the arena must not double the gradient memory.  ResNet-50 (fp32, 161 tensors) and a
    GPT-2-medium-shaped bf16 parameter list: arena <= 1.07 x payload + 8 MiB at world 8 (round 1 reserved
    2 x payload + 1 MiB of LL slots per large parameter)."""
    from b200ddp.reducer import arena_bytes_for

    def fake(numels, dtype):
        return [torch.empty(n, dtype=dtype, device="meta").requires_grad_() for n in numels]

    from torchvision.models import resnet50

    with torch.device("meta"):
        m = resnet50()
    ps = list(m.parameters())
    payload = sum(p.numel() * 4 for p in ps)
    got = arena_bytes_for(ps, world_size=8)
    assert payload < got <= int(1.07 * payload) + (8 << 20)
    # bf16 transformer: 24 x (qkv, proj, fc1, fc2 + norms), tied 50257 x 1024 embedding
    numels = [50257 * 1024, 1024 * 1024]
    for _ in range(24):
        numels += [1024, 1024, 3 * 1024 * 1024, 3 * 1024, 1024 * 1024, 1024, 1024, 1024, 4096 * 1024, 4096, 4096 * 1024, 1024]
    ps = fake(numels, torch.bfloat16)
    payload = sum(numels) * 2
    got = arena_bytes_for(ps, world_size=8)
    assert payload < got <= int(1.07 * payload) + (8 << 20)
    # the toy model: one 84-byte bucket on the LL path: 4 * ws * 96 bytes of slots, nothing like megabytes per tensor
    toy = fake([20, 1], torch.float32)
    assert arena_bytes_for(toy, world_size=8) <= (2 << 20) + 8192
    # small caps put every bucket on the fused (double-buffered) path: two slots each
    small = arena_bytes_for(fake([30000] * 8, torch.float32), world_size=2, bucket_bytes_cap=100_000,
                            first_bucket_bytes_cap=100_000)
    assert small >= 2 * 8 * 30000 * 4


def test_trainer_graph_mode_state_machine_without_a_gpu(monkeypatch, capsys):
    """B200DDP_GRAPH semantics of TrainerCore: `0` is always eager; `auto` runs GRAPH_AFTER eager steps of one
    batch shape, then tries to capture — which cannot work on this CPU box, so the Trainer must report it once
    and keep training eagerly (the same fallback a non-capturable model takes on the GPU)."""
    from b200ddp._trainer_core import TrainerCore

    def make(mode):
        if mode is None:
            monkeypatch.delenv("B200DDP_GRAPH", raising=False)
        else:
            monkeypatch.setenv("B200DDP_GRAPH", mode)
        t = TrainerCore()
        t.model = nn.Linear(4, 1)
        t.optimizer = torch.optim.SGD(t.model.parameters(), lr=0.1)
        t.loss_fn = nn.functional.mse_loss
        t.gpu_id = "cpu"
        return t

    x, y = torch.randn(8, 4), torch.randn(8, 1)
    t = make("0")
    for _ in range(6):
        assert isinstance(t._run_batch(x, y), torch.Tensor)
    assert t._state()["eager_run"] == 0 and t._state()["failed"] is None
    t = make(None)
    assert t._state()["mode"] == "auto"
    w0 = t.model.weight.detach().clone()
    for i in range(6):
        loss = t._run_batch(x, y)
        assert isinstance(loss, torch.Tensor)
        if i == 2:
            assert t._state()["failed"] is None and t._state()["eager_run"] == TrainerCore.GRAPH_AFTER
    assert t._state()["failed"] is not None and not t._state()["slots"]  # capture was attempted once, then eager
    assert "continuing eagerly" in capsys.readouterr().out
    assert not torch.equal(t.model.weight, w0)
    # a different batch shape before capture restarts the count
    t = make(None)
    t._run_batch(x, y)
    t._run_batch(x[:3], y[:3])
    assert t._state()["eager_run"] == 1
    with pytest.raises(ValueError, match="B200DDP_GRAPH"):
        make("sometimes")._state()


def test_comm_hook_has_the_signature_torch_checks():
    """torch's _check_comm_hook (nn/parallel/distributed.py:2268-2289) looks the parameter up by the literal name
    `bucket` and compares annotation OBJECTS."""
    import torch.distributed as dist

    from b200ddp import hooks

    sig = inspect.signature(hooks.allreduce_hook)
    assert list(sig.parameters) == ["state", "bucket"]
    assert sig.parameters["bucket"].annotation is dist.GradBucket
    assert sig.return_annotation == torch.futures.Future[torch.Tensor]


def test_capture_guard_tracks_optimizer_hyperparameters_and_mode():
    """What a captured step has baked in besides shapes (TrainerCore._capture_guard): a learning-rate change (a
    scheduler) or model.eval() must make the Trainer drop its captures."""
    from b200ddp._trainer_core import TrainerCore

    t = TrainerCore()
    t.model = nn.Linear(4, 1)
    t.optimizer = torch.optim.SGD(t.model.parameters(), lr=0.1, momentum=0.9)
    g0 = t._capture_guard()
    assert t._capture_guard() == g0
    t.optimizer.param_groups[0]["lr"] = 0.05
    g1 = t._capture_guard()
    assert g1 != g0
    t.model.eval()
    assert t._capture_guard() != g1
    t.model.train()
    t.optimizer.param_groups[0]["weight_decay"] = 1e-4
    assert t._capture_guard() != g1


def test_deferred_rebuild_moves_live_bucket_views_to_private_storage():
    """gradient_as_bucket_view: the one-time rebuild happens at the NEXT forward; gradients that are still views of
    the old arena layout then (no zero_grad in between: gradient accumulation) must be cloned before the arena is laid
    out anew, otherwise the re-layout would change their meaning."""
    class ViewEngine(RecordingEngine):
        def __init__(self):
            super().__init__()
            self.grad_views = True
            self.arena = torch.zeros(1 << 16)
            self.slots, self.used = {}, 0
            self.cloned_before_plan = None

        def owns(self, t):
            lo = self.arena.data_ptr()
            return t is not None and lo <= t.data_ptr() < lo + self.arena.numel() * 4

        def plan(self, buckets):
            if self.plans:  # the re-layout: no parameter may still point into the arena
                self.cloned_before_plan = [not self.owns(p.grad) for p in self.params]
            super().plan(buckets)

        def reduce_bucket(self, bucket, grads):
            super().reduce_bucket(bucket, grads)
            out = []
            for i, g in zip(bucket.param_indices, grads):  # hand back "views of the arena", like P2PEngine does
                if i not in self.slots:
                    self.slots[i] = self.used
                    self.used += (g.numel() + 15) // 16 * 16
                v = self.arena[self.slots[i]:self.slots[i] + g.numel()].view(g.shape)
                if v.data_ptr() != g.data_ptr():
                    v.copy_(g)
                out.append(v)
            return out

    eng = ViewEngine()
    model = _mlp()
    ddp = DDP(model, bucket_cap_mb=0.01, _engine=eng, gradient_as_bucket_view=True)
    eng.params = list(model.parameters())
    x = torch.randn(8, 20)
    ddp(x).square().mean().backward()
    assert all(eng.owns(p.grad) for p in model.parameters()) and not ddp.reducer.has_rebuilt
    kept = [p.grad.clone() for p in model.parameters()]
    ddp(x).square().mean().backward()  # no zero_grad: the second forward rebuilds with the views still alive
    assert ddp.reducer.has_rebuilt and len(eng.plans) == 2
    assert eng.cloned_before_plan == [True] * 6
    for p, k in zip(model.parameters(), kept):
        assert torch.allclose(p.grad, 2 * k)  # accumulated on top of the preserved values


def test_fuse_optimizer_eligibility_and_per_iteration_arming():
    """DistributedDataParallel.fuse_optimizer (opt-in optimizer epilogue): only plain SGD on exactly this DDP's
    contiguous parameters; armed per iteration only when every gradient starts from None; the wrapped
    optimizer.step is skipped exactly once after an armed backward and steps normally otherwise."""
    class SgdEngine(RecordingEngine):
        def __init__(self):
            super().__init__()
            self.sgd_lr, self.params, self.seen = None, (), []

        def reduce_bucket(self, bucket, grads):
            super().reduce_bucket(bucket, grads)
            self.seen.append(self.sgd_lr)

    # not eligible: momentum, Adam, foreign parameters, an engine without the epilogue
    m = _mlp()
    assert not DDP(m, _engine=SgdEngine()).fuse_optimizer(torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9))
    m = _mlp()
    assert not DDP(m, _engine=SgdEngine()).fuse_optimizer(torch.optim.Adam(m.parameters(), lr=0.1))
    m = _mlp()
    assert not DDP(m, _engine=SgdEngine()).fuse_optimizer(torch.optim.SGD(list(m.parameters())[:2], lr=0.1))
    m = _mlp()
    assert not DDP(m, _engine=RecordingEngine()).fuse_optimizer(torch.optim.SGD(m.parameters(), lr=0.1))

    eng = SgdEngine()
    model = _mlp()
    opt = torch.optim.SGD(model.parameters(), lr=0.25)
    ddp = DDP(model, _engine=eng)
    assert ddp.fuse_optimizer(opt) and ddp.fuse_optimizer(opt)
    x = torch.randn(8, 20)
    before = [p.detach().clone() for p in model.parameters()]
    opt.zero_grad()
    ddp(x).square().mean().backward()
    assert eng.seen and all(lr == 0.25 for lr in eng.seen)  # every bucket was launched with the epilogue
    opt.step()  # the (fake) kernels "applied" the update: the optimizer must not apply it again
    assert all(torch.equal(p, b) for p, b in zip(model.parameters(), before))
    opt.step()  # a second call without a new backward is a normal step
    assert not all(torch.equal(p, b) for p, b in zip(model.parameters(), before))
    # gradient accumulation (no zero_grad): not armed, the optimizer steps itself on the accumulated gradient
    eng.seen.clear()
    opt.param_groups[0]["lr"] = 0.5
    before = [p.detach().clone() for p in model.parameters()]
    ddp(x).square().mean().backward()
    assert eng.seen and all(lr is None for lr in eng.seen)
    opt.step()
    assert not all(torch.equal(p, b) for p, b in zip(model.parameters(), before))
    # armed again after zero_grad, with the new learning rate; no_sync never arms
    eng.seen.clear()
    opt.zero_grad()
    ddp(x).square().mean().backward()
    assert all(lr == 0.5 for lr in eng.seen)
    opt.step()
    eng.seen.clear()
    opt.zero_grad()
    with ddp.no_sync():
        ddp(x).square().mean().backward()
    assert eng.seen == []
    before = [p.detach().clone() for p in model.parameters()]
    opt.step()
    assert not all(torch.equal(p, b) for p, b in zip(model.parameters(), before))
