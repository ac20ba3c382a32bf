"""The C++ hook loop (csrc/native_reducer.cpp, the default on the product engine) exercised on CPU: its C-ABI calls are pointed at recording
ctypes callbacks, so the state machine — post hooks on the AccumulateGrad nodes, pending counts, strict
bucket order, engine-callback finalize, ready-order capture, error paths — is checked without a GPU."""
import ctypes as C

import pytest
import torch
import torch.nn as nn


@pytest.fixture(scope="module")
def native():
    from b200ddp._native import build_native

    return build_native()


class FakeAbi:
    """Recording stand-ins for the entry points the native loop calls (same C signatures as include/b200ddp.h)."""

    def __init__(self):
        self.calls = []
        self.fail_next = None
        vp, u64, i, f = C.c_void_p, C.c_uint64, C.c_int, C.c_float
        pp, pu = C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)

        def rec(name, *keep):
            def cb(*args):
                self.calls.append((name,) + tuple(args[k] for k in keep))
                if self.fail_next == name:
                    self.fail_next = None
                    return -2
                return 0
            return cb

        self._cbs = {
            "b200ddp_stream_wait": C.CFUNCTYPE(i, vp, vp, vp)(rec("stream_wait")),
            "b200ddp_flatten_scale": C.CFUNCTYPE(i, vp, pp, pu, pu, i, i, u64, f, vp)(rec("flatten", 4, 6, 7)),
            "b200ddp_unflatten": C.CFUNCTYPE(i, vp, pp, pu, pu, i, i, u64, vp)(rec("unflatten", 4, 6)),
            "b200ddp_allreduce": C.CFUNCTYPE(i, vp, u64, u64, u64, i, i, i, f, vp)(rec("allreduce", 1, 2, 3, 5, 7)),
            "b200ddp_fused_bucket": C.CFUNCTYPE(i, vp, pp, pu, pu, i, i, u64, u64, f, vp)(rec("fused", 4, 6, 7)),
            "b200ddp_poll_error": C.CFUNCTYPE(i, vp)(rec("poll")),
            "b200ddp_last_error": C.CFUNCTYPE(C.c_char_p, vp)(lambda ctx: b"injected failure"),
        }
        self.table = {k: C.cast(v, C.c_void_p).value for k, v in self._cbs.items()}


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 2))


def _buckets(groups, fused=False, algo=2, views=None):
    out, off = [], 0
    for bi, g in enumerate(groups):
        out.append((g, [16 * k for k in range(len(g))], [4] * len(g), 0, 64 * len(g), off, off, algo, fused,
                    [] if views is None else views[bi]))
        off += 4096
    return out


def test_hooks_launch_buckets_in_index_order_and_finalize(native):
    abi = FakeAbi()
    m = _model()
    params = list(m.parameters())
    red = native.NativeReducer(1234, abi.table, 2, 77, 0, params)
    red.configure(_buckets([[4, 5], [2, 3], [0, 1]]))
    x = torch.randn(3, 6)
    m(x).sum().backward()  # not armed: nothing may happen
    assert abi.calls == [] and red.iterations() == 0
    red.prepare()
    m.zero_grad()
    m(x).sum().backward()
    names = [c[0] for c in abi.calls]
    per_bucket = ["stream_wait", "flatten", "allreduce", "unflatten"]
    assert names == per_bucket * 3 + ["stream_wait", "poll"]
    assert red.launched_buckets() == [0, 1, 2] and red.iterations() == 1 and red.launches() == 9
    flat = [c for c in abi.calls if c[0] == "flatten"]
    assert [c[1] for c in flat] == [2, 2, 2] and [c[2] for c in flat] == [0, 4096, 8192]
    assert all(abs(c[3] - 0.5) < 1e-7 for c in flat)  # 1/world pre-scale
    ar = [c for c in abi.calls if c[0] == "allreduce"]
    assert [(c[1], c[2], c[3], c[4]) for c in ar] == [(0, 0, 128, 2), (4096, 4096, 128, 2), (8192, 8192, 128, 2)]
    assert all(c[5] == 1.0 for c in ar)  # the Reducer pre-scales in flatten; the allreduce itself does not scale
    assert red.take_ready_order() == [5, 4, 3, 2, 1, 0]  # gradient-ready order of a chain
    assert red.take_ready_order() == []


def test_out_of_order_buckets_wait_for_their_predecessors(native):
    abi = FakeAbi()
    m = _model()
    red = native.NativeReducer(1, abi.table, 1, 0, 0, list(m.parameters()))
    # bucket 0 holds the FIRST layer (ready last): nothing may launch before it is complete
    red.configure(_buckets([[0, 1], [2, 3], [4, 5]], fused=True))
    red.prepare()
    m(torch.randn(2, 6)).sum().backward()
    assert [c[0] for c in abi.calls] == ["stream_wait", "fused"] * 3 + ["stream_wait", "poll"]
    assert red.launched_buckets() == [0, 1, 2]
    assert [c[2] for c in abi.calls if c[0] == "fused"] == [0, 4096, 8192]


def test_world1_skips_the_allreduce_and_no_sync_disarms(native):
    abi = FakeAbi()
    m = _model()
    red = native.NativeReducer(1, abi.table, 1, 0, 0, list(m.parameters()))
    red.configure(_buckets([[4, 5], [2, 3], [0, 1]]))
    red.prepare()
    m(torch.randn(2, 6)).sum().backward()
    assert "allreduce" not in [c[0] for c in abi.calls] and red.launches() == 6
    abi.calls.clear()
    red.prepare()
    red.disarm()
    m(torch.randn(2, 6)).sum().backward()
    assert abi.calls == []


def test_errors_surface_as_python_exceptions(native):
    abi = FakeAbi()
    m = _model()
    red = native.NativeReducer(1, abi.table, 2, 0, 0, list(m.parameters()))
    with pytest.raises(RuntimeError, match="not assigned"):
        red.configure(_buckets([[4, 5]]))
    with pytest.raises(RuntimeError, match="two buckets"):
        red.configure(_buckets([[4, 5], [5, 3], [0, 1, 2]]))
    red.configure(_buckets([[4, 5], [2, 3], [0, 1]]))
    red.prepare()
    abi.fail_next = "allreduce"
    with pytest.raises(RuntimeError, match="allreduce failed.*injected failure"):
        m(torch.randn(2, 6)).sum().backward()
    with pytest.raises(RuntimeError, match="missing C-ABI entry point"):
        native.NativeReducer(1, {}, 2, 0, 0, list(m.parameters()))

    class Partial(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(3, 3), nn.Linear(3, 3)

        def forward(self, x):
            return self.a(x)

    p = Partial()
    abi2 = FakeAbi()
    red2 = native.NativeReducer(1, abi2.table, 2, 0, 0, list(p.parameters()))
    red2.configure(_buckets([[2, 3], [0, 1]]))
    red2.prepare()
    with pytest.raises(RuntimeError, match="never produced a gradient"):
        p(torch.randn(2, 3)).sum().backward()
    with pytest.raises(RuntimeError, match="finished reduction in the prior iteration"):
        red2.prepare()


def test_bucket_views_replace_unflatten_and_become_the_grads(native):
    """gradient_as_bucket_view: no unflatten call; .grad of every parameter is re-pointed at the view handed to
    configure(); a second iteration whose gradients ARE the views flattens in place (same pointers)."""
    abi = FakeAbi()
    m = _model()
    params = list(m.parameters())
    groups = [[4, 5], [2, 3], [0, 1]]
    views = [[torch.zeros_like(params[i]) for i in g] for g in groups]
    red = native.NativeReducer(1, abi.table, 2, 0, 0, params)
    red.configure(_buckets(groups, views=views))
    red.prepare()
    m(torch.randn(2, 6)).sum().backward()
    assert [c[0] for c in abi.calls] == ["stream_wait", "flatten", "allreduce"] * 3 + ["stream_wait", "poll"]
    assert red.launches() == 6
    for g, vs in zip(groups, views):
        for i, v in zip(g, vs):
            assert params[i].grad.data_ptr() == v.data_ptr()
    red.rebind(4321)  # a regrown arena: allowed between iterations
    red.prepare()
    m(torch.randn(2, 6)).sum().backward()  # accumulates into the views in place
    for g, vs in zip(groups, views):
        for i, v in zip(g, vs):
            assert params[i].grad.data_ptr() == v.data_ptr()
    with pytest.raises(RuntimeError, match="one view per parameter"):
        red.configure(_buckets(groups, fused=True, views=views))


def test_hooks_are_removed_with_the_reducer(native):
    abi = FakeAbi()
    m = _model()
    red = native.NativeReducer(1, abi.table, 1, 0, 0, list(m.parameters()))
    red.configure(_buckets([[4, 5], [2, 3], [0, 1]]))
    red.prepare()
    red.remove_hooks()
    m(torch.randn(2, 6)).sum().backward()
    assert abi.calls == []
