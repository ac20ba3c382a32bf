"""GPU parity tests proper (run with -m gpu on the B200): every kernel is called through the C ABI and
compared with the oracle — bit-exact for fp32 buckets and for the pure data movement of any dtype.

Multi-rank kernels run here on ONE GPU through a loop-back world (N contexts in this process, peers are
plain pointers): the kernels, signal-pad protocol and summation order are exactly the multi-GPU ones.
Grids are capped so all N ranks' CTAs are co-resident on the one device."""
import numpy as np
import pytest
import torch

from b200ddp import _lib
from b200ddp.comm import P2PComm, TensorTable
from oracle import allreduce_oracle as ao

pytestmark = pytest.mark.gpu

DT = {"f32": (torch.float32, _lib.F32), "bf16": (torch.bfloat16, _lib.BF16), "f16": (torch.float16, _lib.F16)}
ALGOS = {"oneshot": _lib.ALGO_ONESHOT, "twoshot": _lib.ALGO_TWOSHOT, "tma": _lib.ALGO_TWOSHOT_TMA}


def _storage(t: torch.Tensor, dtype: str) -> np.ndarray:
    """Tensor -> the oracle's storage array (float32, or uint16 bit patterns)."""
    t = t.detach().cpu().contiguous()
    if dtype == "f32":
        return t.numpy().ravel().copy()
    return t.view(torch.int16).numpy().view(np.uint16).ravel().copy()


def _rand(n, dtype, gen, device="cuda"):
    return (torch.rand(n, generator=gen, dtype=torch.float32) * 2 - 1).to(device=device, dtype=DT[dtype][0])


@pytest.fixture(scope="module")
def comm1():
    c = P2PComm(0, 1, 0, 64 << 20, _lib.MEM_IPC)
    yield c
    torch.cuda.synchronize()
    c.destroy()


# ---------------------------------------------------------------------------------------------------
# flatten(+scale) / unflatten
# ---------------------------------------------------------------------------------------------------
RAGGED = [1, 3, 4, 5, 1023, 4096, 70001, 2, 0, 8, 2049 * 4 + 1, 20, 1]


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("scale", [1.0, 0.125, 1.0 / 3.0])
def test_flatten_unflatten_bit_exact(comm1, dtype, scale):
    gen = torch.Generator().manual_seed(7)
    grads = [_rand(n, dtype, gen) for n in RAGGED]
    offs = ao.padded_offsets(RAGGED, dtype)
    epv = 16 // ao.elem_size(dtype)
    total = offs[-1] + (RAGGED[-1] + epv - 1) // epv * epv
    s = torch.cuda.current_stream().cuda_stream
    comm1.arena().zero_()
    table = TensorTable([g.data_ptr() for g in grads], offs, RAGGED)
    comm1.flatten_scale(table, DT[dtype][1], 4096, scale, s)
    torch.cuda.synchronize()
    got = _storage(comm1.view(4096, total * ao.elem_size(dtype), DT[dtype][0]), dtype)
    want = ao.flatten_scale([_storage(g, dtype) for g in grads], offs, total, scale, dtype)
    assert np.array_equal(got, want)
    # unflatten into fresh tensors: pure copy of each tensor's window
    outs = [torch.full_like(g, 7.0) for g in grads]
    comm1.unflatten(TensorTable([o.data_ptr() for o in outs], offs, RAGGED), DT[dtype][1], 4096, s)
    torch.cuda.synchronize()
    for o, off, n in zip(outs, offs, RAGGED):
        assert np.array_equal(_storage(o, dtype), want[off:off + n])
    comm1.poll_error()


def test_scale_one_is_bitwise_passthrough(comm1):
    """scale == 1 moves raw bytes (parameter/buffer broadcast of any dtype): NaN payloads, signalling NaNs,
    denormals and reinterpreted integers must survive flatten and unflatten bit for bit."""
    pats = torch.tensor([-1, 0x7F800001, 0x7FC12345, 0x00000001, -2147483648, 0x7F7FFFFF, 1234567, 0, 5],
                        dtype=torch.int64).to(torch.int32).cuda()
    big = torch.randint(-2**31, 2**31 - 1, (4096 * 3 + 1,), dtype=torch.int64).to(torch.int32).cuda()
    srcs = [pats, big]
    lens = [t.numel() for t in srcs]
    offs = ao.padded_offsets(lens, "f32")
    s = torch.cuda.current_stream().cuda_stream
    comm1.flatten_scale(TensorTable([t.data_ptr() for t in srcs], offs, lens), _lib.F32, 0, 1.0, s)
    outs = [torch.zeros_like(t) for t in srcs]
    comm1.unflatten(TensorTable([t.data_ptr() for t in outs], offs, lens), _lib.F32, 0, s)
    torch.cuda.synchronize()
    for a, b in zip(srcs, outs):
        assert torch.equal(a, b)
    # 16-bit raw path
    h = torch.randint(-32768, 32767, (1001,), dtype=torch.int16).cuda()
    o = torch.zeros_like(h)
    comm1.flatten_scale(TensorTable([h.data_ptr()], [0], [1001]), _lib.BF16, 0, 1.0, s)
    comm1.unflatten(TensorTable([o.data_ptr()], [0], [1001]), _lib.BF16, 0, s)
    torch.cuda.synchronize()
    assert torch.equal(h, o)


def test_flatten_unaligned_sources_and_many_tiny_tensors(comm1):
    """Storage offsets that break 16-byte alignment, and > one table (160) of tiny tensors (ResNet-50 has
    107 gradients below 16 KiB)."""
    gen = torch.Generator().manual_seed(3)
    base = _rand(600 * 67 + 5, "f32", gen)
    lens = [(i % 61) + 1 for i in range(500)]
    grads, pos = [], 1  # start at element 1: 4-byte aligned only
    for n in lens:
        grads.append(base[pos:pos + n])
        pos += n + (i_pad := (n % 3))
    assert any(g.data_ptr() % 16 for g in grads)
    offs = ao.padded_offsets(lens, "f32")
    total = offs[-1] + (lens[-1] + 3) // 4 * 4
    s = torch.cuda.current_stream().cuda_stream
    n0 = comm1.launch_count
    comm1.flatten_scale(TensorTable([g.data_ptr() for g in grads], offs, lens), _lib.F32, 0, 0.5, s)
    assert comm1.launch_count - n0 == 4  # ceil(500 / 160) table chunks
    torch.cuda.synchronize()
    want = ao.flatten_scale([_storage(g, "f32") for g in grads], offs, total, 0.5, "f32")
    assert np.array_equal(_storage(comm1.view(0, total * 4, torch.float32), "f32"), want)
    outs_base = torch.zeros_like(base)
    outs, pos = [], 1
    for n in lens:
        outs.append(outs_base[pos:pos + n])
        pos += n + (n % 3)
    comm1.unflatten(TensorTable([o.data_ptr() for o in outs], offs, lens), _lib.F32, 0, s)
    torch.cuda.synchronize()
    for o, off, n in zip(outs, offs, lens):
        assert np.array_equal(_storage(o, "f32"), want[off:off + n])
    # gaps between the scattered windows were not touched
    mask = torch.ones_like(outs_base, dtype=torch.bool)
    pos = 1
    for n in lens:
        mask[pos:pos + n] = False
        pos += n + (n % 3)
    assert float(outs_base[mask].abs().max()) == 0.0


def test_large_bucket_and_argument_validation(comm1):
    gen = torch.Generator().manual_seed(5)
    lens = [4096 * 1024 + 3, 17, 2 * 1024 * 1024]  # ~24 MiB: fast path + tails
    grads = [_rand(n, "f32", gen) for n in lens]
    offs = ao.padded_offsets(lens, "f32")
    total = offs[-1] + (lens[-1] + 3) // 4 * 4
    s = torch.cuda.current_stream().cuda_stream
    comm1.flatten_scale(TensorTable([g.data_ptr() for g in grads], offs, lens), _lib.F32, 0, 0.25, s)
    torch.cuda.synchronize()
    want = ao.flatten_scale([_storage(g, "f32") for g in grads], offs, total, 0.25, "f32")
    assert np.array_equal(_storage(comm1.view(0, total * 4, torch.float32), "f32"), want)
    # errors: misaligned bucket offset, overlapping tensors, out-of-range, unknown dtype
    t = TensorTable([grads[1].data_ptr()], [0], [17])
    with pytest.raises(RuntimeError, match="16"):
        comm1.flatten_scale(t, _lib.F32, 8, 1.0, s)
    with pytest.raises(RuntimeError, match="overlap"):
        comm1.flatten_scale(TensorTable([grads[1].data_ptr()] * 2, [0, 4], [17, 17]), _lib.F32, 0, 1.0, s)
    with pytest.raises(RuntimeError, match="exceeds"):
        comm1.flatten_scale(t, _lib.F32, comm1.arena_bytes - 16, 1.0, s)
    with pytest.raises(RuntimeError, match="dtype"):
        comm1.flatten_scale(t, 9, 0, 1.0, s)
    with pytest.raises(RuntimeError, match="not 16-byte aligned"):
        comm1.flatten_scale(TensorTable([grads[1].data_ptr()], [2], [17]), _lib.F32, 0, 1.0, s)


def test_world1_fused_bucket_is_identity_scale(comm1):
    gen = torch.Generator().manual_seed(11)
    lens = [20, 1]  # the toy model's bucket: 84 B
    grads = [_rand(n, "f32", gen) for n in lens]
    before = [g.clone() for g in grads]
    offs = ao.padded_offsets(lens, "f32")
    s = torch.cuda.current_stream().cuda_stream
    comm1.fused_bucket(TensorTable([g.data_ptr() for g in grads], offs, lens), _lib.F32, 0, 96, 1.0, s)
    torch.cuda.synchronize()
    for g, b in zip(grads, before):
        assert torch.equal(g, b)
    assert comm1.pick_algo(1 << 20) in ALGOS.values() or comm1.pick_algo(1 << 20) == _lib.ALGO_NVLS


# ---------------------------------------------------------------------------------------------------
# allreduce kernels on loop-back worlds
# ---------------------------------------------------------------------------------------------------
class World:
    def __init__(self, ws, arena=48 << 20):
        self.ws = ws
        self.comms = P2PComm.loopback_world(ws, 0, arena)
        self.streams = [torch.cuda.Stream() for _ in range(ws)]
        for c in self.comms:
            c.set_timeout_ms(5000)

    def close(self):
        torch.cuda.synchronize()
        for c in self.comms:
            c.destroy()

    def fill(self, off, n_el, dtype, seed):
        bufs = []
        for r, c in enumerate(self.comms):
            gen = torch.Generator().manual_seed(seed * 100 + r)
            t = _rand(n_el, dtype, gen)
            c.view(off, n_el * ao.elem_size(dtype), DT[dtype][0]).copy_(t)
            bufs.append(_storage(t, dtype))
        torch.cuda.synchronize()
        return bufs

    def run(self, fn):
        """fn(rank, comm, stream_ptr) enqueues on each rank's own stream; then everything is synchronised."""
        for r, (c, st) in enumerate(zip(self.comms, self.streams)):
            fn(r, c, st.cuda_stream)
        torch.cuda.synchronize()
        for c in self.comms:
            c.poll_error()


@pytest.fixture(scope="module", params=[2, 3, 4, 8])
def world(request):
    w = World(request.param)
    yield w
    w.close()


@pytest.mark.parametrize("algo", ["oneshot", "twoshot", "tma"])
@pytest.mark.parametrize("dtype,n_el", [("f32", 4), ("f32", 4 * 1000 + 4), ("f32", 4 * 65536 + 12), ("bf16", 8 * 5001),
                                        ("f16", 8 * 777), ("f32", 1 << 20)])
def test_allreduce_bit_exact_vs_oracle(world, algo, dtype, n_el):
    nbytes = n_el * ao.elem_size(dtype)
    in_off, out_off = 4096, (4096 + nbytes + 4095) // 4096 * 4096 + 4096
    if algo != "oneshot" and n_el % 8 == 0:
        out_off = in_off  # exercise the in-place form on some cases
    bufs = world.fill(in_off, n_el, dtype, seed=n_el % 97)
    want = ao.allreduce_sum(bufs, dtype)
    world.run(lambda r, c, s: c.allreduce(in_off, out_off, nbytes, DT[dtype][1], ALGOS[algo], 8, s))
    for c in world.comms:
        got = _storage(c.view(out_off, nbytes, DT[dtype][0]), dtype)
        assert np.array_equal(got, want), f"{algo} ws={world.ws} rank {c.rank}"


def test_allreduce_repeated_launches_and_grid_sizes(world):
    """Epochs advance per CTA in device memory: back-to-back launches with different grids must stay in
    step (this is what makes the launch CUDA-graph replayable)."""
    n_el = 4 * 30000
    nbytes = n_el * 4
    for i, (algo, ctas) in enumerate([("twoshot", 2), ("oneshot", 7), ("tma", 3), ("twoshot", 12), ("tma", 1)]):
        bufs = world.fill(0, n_el, "f32", seed=i)
        want = ao.allreduce_sum(bufs, "f32")
        out = 1 << 20
        world.run(lambda r, c, s: c.allreduce(0, out, nbytes, _lib.F32, ALGOS[algo], ctas, s))
        for c in world.comms:
            assert np.array_equal(_storage(c.view(out, nbytes, torch.float32), "f32"), want)


def test_linearity_at_full_bucket_size(world):
    """Size-independent property at a full 25 MiB bucket: allreduce(a) + allreduce(b) == allreduce(a + b)
    for integer-valued fp32 data (exact arithmetic), and the result equals ws * value for constants."""
    n_el = (25 << 20) // 4
    nbytes = n_el * 4
    for c in world.comms:
        c.view(0, nbytes, torch.float32).copy_(torch.arange(n_el, device="cuda", dtype=torch.float32) % 1024 + c.rank)
    torch.cuda.synchronize()
    world.run(lambda r, c, s: c.allreduce(0, 0, nbytes, _lib.F32, _lib.ALGO_TWOSHOT, 16, s))
    ws = world.ws
    expect = (torch.arange(n_el, device="cuda", dtype=torch.float32) % 1024) * ws + ws * (ws - 1) / 2
    for c in world.comms:
        assert torch.equal(c.view(0, nbytes, torch.float32), expect)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("lens", [[20, 1], [20, 1, 300, 7, 4096, 1], [3, 9001, 5, 20000, 2]],
                         ids=["toy84B-LL", "17KB-LL", "116KB-barrier"])
def test_fused_bucket_matches_three_step_oracle(world, dtype, lens):
    offs = ao.padded_offsets(lens, dtype)
    epv = 16 // ao.elem_size(dtype)
    total = offs[-1] + epv
    nbytes = total * ao.elem_size(dtype)
    ws = world.ws
    scale = float(np.float32(1.0) / np.float32(ws))
    grads = []
    for r in range(ws):
        gen = torch.Generator().manual_seed(50 + r)
        grads.append([_rand(n, dtype, gen) for n in lens])
    want = ao.ddp_average([[_storage(g, dtype) for g in gs] for gs in grads], dtype)
    region = world.comms[0].fused_region_bytes(nbytes)
    assert region == (4 * ws * nbytes if nbytes <= 32768 else 2 * nbytes)
    world.run(lambda r, c, s: c.clear(1 << 20, region, s))  # tagged slots start from zero
    world.run(lambda r, c, s: c.fused_bucket(TensorTable([g.data_ptr() for g in grads[r]], offs, lens),
                                             DT[dtype][1], 1 << 20, nbytes, scale, s))
    for r in range(ws):
        for g, w in zip(grads[r], want):
            assert np.array_equal(_storage(g, dtype), w)


@pytest.mark.parametrize("lens", [[20, 1, 1000, 3], [20, 1, 9000, 3]], ids=["LL", "barrier"])
def test_fused_bucket_repeated_calls_alternate_staging_slots(world, lens):
    """The fused kernels have at most ONE barrier (none on the LL path) and double-buffered staging:
    back-to-back calls on the same bucket (different data every time, no host sync in between) must never
    read a stale or half-written slot."""
    offs = ao.padded_offsets(lens, "f32")
    total = offs[-1] + 4
    ws = world.ws
    scale = float(np.float32(1.0) / np.float32(ws))
    rounds = 6
    grads = [[[_rand(n, "f32", torch.Generator().manual_seed(1000 * k + 10 * r + i)) for i, n in enumerate(lens)]
              for r in range(ws)] for k in range(rounds)]
    wants = [ao.ddp_average([[_storage(g, "f32") for g in gs] for gs in grads[k]]) for k in range(rounds)]
    world.run(lambda r, c, s: c.clear(16384, c.fused_region_bytes(total * 4), s))
    for k in range(rounds):  # all rounds enqueued before any synchronisation
        for r, (c, st) in enumerate(zip(world.comms, world.streams)):
            c.fused_bucket(TensorTable([g.data_ptr() for g in grads[k][r]], offs, lens), _lib.F32, 16384, total * 4,
                           scale, st.cuda_stream)
    torch.cuda.synchronize()
    for c in world.comms:
        c.poll_error()
    for k in range(rounds):
        for r in range(ws):
            for g, w in zip(grads[k][r], wants[k]):
                assert np.array_equal(_storage(g, "f32"), w), f"round {k} rank {r}"


def test_broadcast_and_barrier(world):
    n = 4 * 100003
    for root in (0, world.ws - 1):
        bufs = world.fill(0, n, "f32", seed=root + 1)
        world.run(lambda r, c, s: c.broadcast(0, n * 4, root, s, 4))
        for c in world.comms:
            assert np.array_equal(_storage(c.view(0, n * 4, torch.float32), "f32"), bufs[root])
    world.run(lambda r, c, s: c.barrier(s))


def test_missing_peer_times_out_instead_of_hanging():
    w = World(2, arena=1 << 20)
    try:
        for c in w.comms:
            c.set_timeout_ms(200)
        st = torch.cuda.Stream()
        w.comms[0].allreduce(0, 4096, 1024, _lib.F32, _lib.ALGO_ONESHOT, 1, st.cuda_stream)  # rank 1 never launches
        torch.cuda.synchronize()
        with pytest.raises(RuntimeError, match="timed out"):
            w.comms[0].poll_error()
    finally:
        w.close()


def test_timeout_poisons_the_rank_and_gradients_are_left_alone():
    """after a peer timeout a kernel must not scatter a partial sum.  Rank 1 never launches: rank 0's
    allreduce times out, the unflatten that follows it leaves the gradient untouched, the rank stays poisoned
    (later launches return at once instead of waiting for another timeout), and the host sees the error."""
    import time

    w = World(2, arena=1 << 20)
    try:
        c0 = w.comms[0]
        c0.set_timeout_ms(200)
        s = torch.cuda.Stream().cuda_stream
        gen = torch.Generator().manual_seed(5)
        g = _rand(1000, "f32", gen)
        before = g.clone()
        table = TensorTable([g.data_ptr()], [0], [1000])
        c0.flatten_scale(table, _lib.F32, 0, 0.5, s)
        c0.allreduce(0, 0, 4000, _lib.F32, _lib.ALGO_TWOSHOT, 1, s)
        c0.unflatten(table, _lib.F32, 0, s)
        torch.cuda.synchronize()
        assert torch.equal(g, before), "unflatten scattered a bucket whose allreduce had timed out"
        assert torch.equal(c0.view(0, 4000, torch.float32), before * 0.5)  # the flatten itself did run
        small = [_rand(20, "f32", gen), _rand(1, "f32", gen)]
        keep = [t.clone() for t in small]
        c0.clear(8192, c0.fused_region_bytes(96), s)
        t0 = time.perf_counter()
        c0.fused_bucket(TensorTable([t.data_ptr() for t in small], [0, 20], [20, 1]), _lib.F32, 8192, 96, 0.5, s)
        c0.allreduce(0, 0, 4000, _lib.F32, _lib.ALGO_TWOSHOT_TMA, 1, s)
        torch.cuda.synchronize()
        assert time.perf_counter() - t0 < 0.15, "a poisoned rank must not wait for another timeout"
        for t, k in zip(small, keep):
            assert torch.equal(t, k)
        with pytest.raises(RuntimeError, match="timed out"):
            c0.poll_error()
    finally:
        w.close()


def test_allreduce_scale_is_applied_to_the_reduced_sum(world):
    """`scale` of b200ddp_allreduce (the comm-hook shim's 1/world): out = fl32(scale * sum), one extra rounding."""
    n = 4 * 3001
    bufs = world.fill(0, n, "f32", seed=9)
    want = ao.allreduce_sum(bufs, "f32") * np.float32(0.25)
    for name, algo in ALGOS.items():
        world.fill(0, n, "f32", seed=9)
        out_off = (1 << 20) if name == "oneshot" else 0
        world.run(lambda r, c, s: c.allreduce(0, out_off, n * 4, _lib.F32, algo, 4, s, scale=0.25))
        for c in world.comms:
            assert np.array_equal(_storage(c.view(out_off, n * 4, torch.float32), "f32"), want), name


def test_world1_allreduce_with_scale_scales_in_place(comm1):
    gen = torch.Generator().manual_seed(2)
    t = _rand(4 * 777, "f32", gen)
    comm1.view(0, t.numel() * 4, torch.float32).copy_(t)
    s = torch.cuda.current_stream().cuda_stream
    comm1.allreduce(0, 0, t.numel() * 4, _lib.F32, _lib.ALGO_AUTO, 0, s, scale=0.125)
    torch.cuda.synchronize()
    assert torch.equal(comm1.view(0, t.numel() * 4, torch.float32), t * 0.125)


def test_peer_read_copies_a_peers_arena(world):
    n = 4 * 5003
    bufs = world.fill(0, n, "f32", seed=4)
    peer = (0 + 1) % world.ws
    c0, s0 = world.comms[0], world.streams[0].cuda_stream
    c0.peer_read(peer, 0, 1 << 20, n * 4, s0, 4)
    torch.cuda.synchronize()
    assert np.array_equal(_storage(c0.view(1 << 20, n * 4, torch.float32), "f32"), bufs[peer])


def test_allreduce_argument_validation(world):
    c = world.comms[0]
    s = world.streams[0].cuda_stream
    with pytest.raises(RuntimeError, match="out of place"):
        c.allreduce(0, 0, 1024, _lib.F32, _lib.ALGO_ONESHOT, 1, s)
    with pytest.raises(RuntimeError, match="multiples of 16"):
        c.allreduce(0, 4096, 1000, _lib.F32, _lib.ALGO_TWOSHOT, 1, s)
    with pytest.raises(RuntimeError, match="overlap"):
        c.allreduce(0, 512, 1024, _lib.F32, _lib.ALGO_TWOSHOT, 1, s)
    with pytest.raises(RuntimeError, match="multicast"):
        c.allreduce(0, 0, 1024, _lib.F32, _lib.ALGO_NVLS, 1, s)


def test_gather_rows_follows_the_cursor_and_wraps():
    """b200ddp_gather_rows: batch b of a resident shard = rows (cursor + i) mod full, cursor advanced by
    the kernel itself; with a split the source and target columns land in two dense tensors."""
    import ctypes as C

    lib = _lib.load()
    gen = torch.Generator().manual_seed(13)
    full, width, fx, bsz = 100, 21, 20, 32
    shard = torch.rand(full, width, generator=gen).cuda()
    cursor = torch.zeros((), dtype=torch.int64, device="cuda")
    whole = torch.empty(bsz, width, device="cuda")
    bx, by = torch.empty(bsz, fx, device="cuda"), torch.empty(bsz, width - fx, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    pos = 0
    for step in range(9):  # 9 * 32 rows: wraps around the 100-row shard twice
        split = step % 2 == 1
        rc = lib.b200ddp_gather_rows(shard.data_ptr(), bx.data_ptr() if split else whole.data_ptr(),
                                     by.data_ptr() if split else None, width * 4, fx * 4 if split else 0, bsz, full,
                                     cursor.data_ptr(), s)
        assert rc == 0, _lib.last_error(None)
        torch.cuda.synchronize()
        idx = (torch.arange(bsz) + pos) % full
        want = shard[idx.cuda()]
        if split:
            assert torch.equal(bx, want[:, :fx]) and torch.equal(by, want[:, fx:])
        else:
            assert torch.equal(whole, want)
        pos = (pos + bsz) % full
        assert int(cursor) == pos
    assert lib.b200ddp_gather_rows(shard.data_ptr(), whole.data_ptr(), None, 6, 0, bsz, full, cursor.data_ptr(), s) == -1
    assert "4-byte words" in _lib.last_error(None)


@pytest.mark.parametrize("lens", [[20, 1], [30000, 5, 4099, 12001]], ids=["ll-84B", "one-barrier-184KB"])
def test_fused_bucket_sgd_epilogue_on_loopback_worlds(world, lens):
    """b200ddp_fused_bucket_sgd at ws in {2,3,4,8}: gradients = the oracle's average (bit-exact, as without the
    epilogue) and parameters = fma(-lr, g_avg, p) on every rank (one rounding: compared in float64, <= 1 ulp)."""
    ws = world.ws
    lr = 0.125
    offs = ao.padded_offsets(lens, "f32")
    total = offs[-1] + (lens[-1] + 3) // 4 * 4
    gens = [torch.Generator().manual_seed(40 + r) for r in range(ws)]
    grads = [[_rand(n, "f32", g) for n in lens] for g in gens]
    pgen = torch.Generator().manual_seed(3)
    p0 = [_rand(n, "f32", pgen) for n in lens]
    params = [[p.clone() for p in p0] for _ in range(ws)]
    want_g = ao.ddp_average([[_storage(g, "f32") for g in gs] for gs in grads])
    region = world.comms[0].fused_region_bytes(total * 4)
    world.run(lambda r, c, s: c.clear(1 << 20, (region + 15) // 16 * 16, s))

    def launch(r, c, s):
        table = TensorTable([g.data_ptr() for g in grads[r]], offs, lens, [p.data_ptr() for p in params[r]])
        c.fused_bucket_sgd(table, _lib.F32, 1 << 20, total * 4, 1.0 / ws, lr, s)

    world.run(launch)
    for r in range(ws):
        for g, w, p, q in zip(grads[r], want_g, params[r], p0):
            assert np.array_equal(_storage(g, "f32"), w)
            want_p = (q.cpu().double() - lr * torch.from_numpy(w).double()).float()
            err = (p.cpu().double() - want_p.double()).abs().max() / max(1e-30, float(want_p.abs().max()))
            assert float(err) <= 2e-7, f"rank {r}: parameter update off by {float(err):.3e}"
