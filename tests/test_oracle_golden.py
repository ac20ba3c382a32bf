"""Pins the oracle (and the product's host-side bucketing) against the golden fixtures recorded from the
reference stack itself (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from b200ddp.bucketing import compute_bucket_assignment_by_size
from oracle import allreduce_oracle as ao
from oracle import bucket_oracle as bo
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ESIZE = {"f32": 4, "bf16": 2, "f16": 2}


def _bucket_cases():
    with open(os.path.join(GOLD, "bucket_assignment.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _bucket_cases(), ids=lambda c: c["name"])
def test_bucket_assignment_oracle_and_product_match_torch(case):
    nbytes = [n * ESIZE[d] for n, d in zip(case["numels"], case["dtypes"])]
    keys = [(d, "dev0") for d in case["dtypes"]]
    got_o, lim_o = bo.assign_buckets(nbytes, keys, case["limits"], case["tensor_indices"])
    got_p, lim_p = compute_bucket_assignment_by_size(nbytes, keys, case["limits"], case["tensor_indices"])
    if case["tensor_indices"] is not None and len(set(case["dtypes"])) > 1:
        # rule R5: with several (dtype, device) keys AND ready-order input, torch emits the left-over
        # accumulators in std::unordered_map iteration order (implementation defined); membership and
        # limits are pinned, the order of those trailing buckets is ours (first-seen key)
        def norm(b, l):
            return sorted(zip(map(tuple, b), l))
        assert norm(got_o, lim_o) == norm(case["expect_buckets"], case["expect_limits"])
        assert norm(got_p, lim_p) == norm(case["expect_buckets"], case["expect_limits"])
        assert got_o == got_p
        return
    assert got_o == case["expect_buckets"]
    assert lim_o == case["expect_limits"]
    assert got_p == case["expect_buckets"]
    assert lim_p == case["expect_limits"]


def test_bucket_oracle_edge_cases():
    assert bo.assign_buckets([], [], [1 << 20]) == ([], [])
    assert compute_bucket_assignment_by_size([], [], [1 << 20]) == ([], [])
    # zero-sized tensors never close a bucket on their own
    assert bo.assign_buckets([0, 0, 8], ["k"] * 3, [8])[0] == [[0, 1, 2]]
    assert compute_bucket_assignment_by_size([0, 0, 8], ["k"] * 3, [8])[0] == [[0, 1, 2]]


def test_ddp_helpers_reverse_and_rebuild():
    nbytes = [4 * n for n in (300000, 10, 70000, 200000)]
    keys = ["k"] * 4
    assert bo.ddp_initial_buckets(nbytes, keys) == [[1, 2, 3], [0]]
    assert bo.ddp_rebuilt_buckets(nbytes, keys, [3, 2, 1, 0]) == [[3, 2], [1, 0]]


def test_closed_form_matches_reference_trainer_ws2():
    """The reference's own Trainer (gloo, ws=2): sum_r g_r * fl32(1/ws) in rank order == DDP's .grad."""
    fx = torch.load(os.path.join(GOLD, "ref_trainer_ws2.pt"))
    mse = fx["cases"]["mse"]
    assert len(mse[0]["ddp_grads"]) == 4
    for step in range(4):
        per_rank = [[g.numpy() for g in mse[r]["local_grads"][step]] for r in range(2)]
        avg = ao.ddp_average(per_rank)
        for r in range(2):
            for a, ref in zip(avg, mse[r]["ddp_grads"][step]):
                assert float(ref.abs().max()) > 0  # non-vacuous
                assert rel_err(torch.from_numpy(a), ref) <= 1e-6
                # at ws=2 there is only one possible summation order: bitwise equal
                assert np.array_equal(a, ref.numpy().ravel())
    # SGD step the reference took with those gradients (lr=1e-3, ref:multigpu.py:68)
    w0, g, w1 = mse[0]["w_before"][0], mse[0]["ddp_grads"][0], mse[0]["w_after"][0]
    for a, b, c in zip(w0, g, w1):
        assert torch.allclose(a - 1e-3 * b, c, rtol=0, atol=1e-7)
    # the scripts' CrossEntropyLoss on one logit: gradients identically zero
    ce = fx["cases"]["ce"]
    assert all(float(t.abs().max()) == 0.0 for t in ce[0]["ddp_grads"][0])
    # both ranks hold the same weights after DDP's construction-time broadcast
    for a, b in zip(mse[0]["w_before"][0], mse[1]["w_before"][0]):
        assert torch.equal(a, b)
    # DistributedSampler: rank r sees perm[r::2], different rows per rank
    assert mse[0]["sampler_indices"] != mse[1]["sampler_indices"]


def test_closed_form_matches_stock_ddp_ws3():
    fx = torch.load(os.path.join(GOLD, "stock_ddp_mlp_ws3.pt"))
    ranks = fx["ranks"]
    for step in range(2):
        per_rank = [[g.numpy() for g in ranks[r]["local_grads"][step]] for r in range(3)]
        avg = ao.ddp_average(per_rank)
        for a, ref in zip(avg, ranks[0]["ddp_grads"][step]):
            assert rel_err(torch.from_numpy(a), ref) <= 1e-6


@pytest.mark.parametrize("ws", [1, 2, 3, 8])
def test_c_restatement_equals_numpy_restatement(ws):
    import ctypes as C

    from oracle.build import load

    lib = load()
    rng = np.random.default_rng(ws)
    lengths = [1, 7, 4096, 33, 20, 1]
    offs = ao.padded_offsets(lengths, "f32")
    total = offs[-1] + 4
    grads = [[rng.uniform(-1, 1, n).astype(np.float32) for n in lengths] for _ in range(ws)]
    want = ao.ddp_average(grads)
    pf = C.POINTER(C.c_float)
    work = [[g.copy() for g in gs] for gs in grads]
    rank_ptrs = [(pf * len(lengths))(*[g.ctypes.data_as(pf) for g in gs]) for gs in work]
    all_ptrs = (C.POINTER(pf) * ws)(*[C.cast(p, C.POINTER(pf)) for p in rank_ptrs])
    off_a = (C.c_uint64 * len(lengths))(*offs)
    len_a = (C.c_uint64 * len(lengths))(*lengths)
    scratch = np.zeros(ws * total, dtype=np.float32)
    nbytes = lib.oracle_ddp_step(all_ptrs, off_a, len_a, len(lengths), ws, total, scratch.ctypes.data_as(pf))
    assert nbytes == 4 * sum(lengths)
    for r in range(ws):
        for got, exp in zip(work[r], want):
            assert np.array_equal(got, exp)


def test_bf16_helpers_match_torch():
    x = torch.randn(10000, dtype=torch.float32) * 3
    x[:4] = torch.tensor([0.0, -0.0, float("inf"), 1e-40])
    bits = ao.f32_to_bf16_bits(x.numpy())
    ref = x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(bits, ref)
    back = ao.bf16_bits_to_f32(bits)
    assert np.array_equal(back, x.to(torch.bfloat16).float().numpy())


def test_bf16_average_rounds_once():
    rng = np.random.default_rng(0)
    grads = [[ao.f32_to_bf16_bits(rng.normal(size=100).astype(np.float32))] for _ in range(4)]
    out = ao.ddp_average(grads, "bf16")[0]
    # reference computation in float64 of the same pre-scaled bf16 values
    scaled = [ao.bf16_bits_to_f32(ao.f32_to_bf16_bits(ao.bf16_bits_to_f32(g[0]) * np.float32(0.25))) for g in grads]
    exact = np.sum(np.stack(scaled).astype(np.float64), axis=0)
    got = ao.bf16_bits_to_f32(out).astype(np.float64)
    assert np.max(np.abs(got - exact) / np.maximum(np.abs(exact), 1e-3)) < 2 ** -7


def test_round2_oracle_extras_are_pinned_against_torch():
    """`scale` of the allreduce and the optimizer epilogue, restated in the oracle, against torch on CPU: the scaled
    sum is the rank-ordered fp32 sum times fl32(scale); the epilogue is torch.optim.SGD's `param.add_(grad, alpha=-lr)`
    within one fp32 ulp (torch's CPU kernel may or may not contract the multiply-add), bf16 within one bf16 ulp."""
    import numpy as np
    import torch

    from oracle import allreduce_oracle as ao

    rng = np.random.default_rng(7)
    bufs = [rng.uniform(-1, 1, 4099).astype(np.float32) for _ in range(5)]
    plain = ao.allreduce_sum(bufs, "f32")
    assert np.array_equal(ao.allreduce_sum_scaled(bufs, "f32", 1.0), plain)
    assert np.array_equal(ao.allreduce_sum_scaled(bufs, "f32", 0.2), plain * np.float32(0.2))
    acc = torch.zeros(4099)
    acc.copy_(torch.from_numpy(bufs[0]))
    for b in bufs[1:]:
        acc = acc + torch.from_numpy(b)
    assert np.array_equal((acc * 0.2).numpy(), ao.allreduce_sum_scaled(bufs, "f32", 0.2))

    p = torch.from_numpy(rng.uniform(-1, 1, 4099).astype(np.float32))
    g = torch.from_numpy(rng.uniform(-1, 1, 4099).astype(np.float32))
    param = torch.nn.Parameter(p.clone())
    param.grad = g.clone()
    torch.optim.SGD([param], lr=0.0375).step()
    want = param.detach().numpy()
    got = ao.sgd_epilogue(p.numpy(), g.numpy(), 0.0375)
    ulp = np.spacing(np.abs(want).astype(np.float32))
    assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= ulp)
    # bf16 storage
    pb, gb = p.bfloat16(), g.bfloat16()
    param = torch.nn.Parameter(pb.clone())
    param.grad = gb.clone()
    torch.optim.SGD([param], lr=0.0375).step()
    bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    got = ao.bf16_bits_to_f32(ao.sgd_epilogue(bits(pb), bits(gb), 0.0375, "bf16"))
    want = param.detach().float().numpy()
    scale = np.maximum(np.abs(want), np.abs(pb.float().numpy()))  # a result near zero is a cancellation: one ulp of p
    assert np.all(np.abs(got - want) <= scale * 2.0 ** -7 + 1e-30)
