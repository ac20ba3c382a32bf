"""Custom gradient Reducer — the host side of the hot path.

Mirrors the behaviour of ``c10d::Reducer`` as the reference's ``DDP(model)`` uses it
(torch:include/torch/csrc/distributed/c10d/reducer.hpp): a post-accumulate hook per parameter
(``autograd_hook``, :74) counts a bucket's pending gradients down (``mark_variable_ready``, :280); a
bucket whose count reaches zero is launched as soon as all earlier buckets have been
(``mark_bucket_ready``, :282, strict index order via ``next_bucket_``, :231); a callback queued on the
autograd engine ends the iteration (``finalize_backward``, :286); after the first iteration the buckets
are rebuilt once in the order gradients actually became ready (``rebuild_buckets``, :136,149-151) with
rank 0's order made authoritative (``sync_bucket_indices``, :299-301).

What differs is below the bucket: instead of one D2D copy per parameter into the bucket, an NCCL
allreduce, and one copy per parameter back, a ready bucket is handed to a ``BucketEngine`` whose product
implementation (``P2PEngine``) enqueues hand-written sm_100a kernels through the C ABI: one multi-tensor
flatten(+1/ws scale), one peer-to-peer allreduce, one multi-tensor unflatten — or a single fused launch
for small buckets.  There is no CPU or NCCL implementation of the engine in this package.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
from torch.autograd import Variable

from . import _lib, _native
from .bucketing import (
    DEFAULT_BUCKET_BYTES_CAP,
    DEFAULT_FIRST_BUCKET_BYTES,
    BucketLayout,
    BucketSpec,
    compute_bucket_assignment_by_size,
    layout_bucket,
)
from .comm import P2PComm, TensorTable, dtype_code


class BucketEngine:
    """What the Reducer needs from the layer that moves bytes.  ``P2PEngine`` is the only implementation
    shipped; tests inject recording/gloo engines to exercise the host logic without a GPU."""

    rank: int = 0
    world_size: int = 1

    def plan(self, buckets: List[BucketSpec]) -> None:
        """Called whenever the bucket list changes (construction, rebuild)."""
        raise NotImplementedError

    def reduce_bucket(self, bucket: BucketSpec, grads: List[torch.Tensor]) -> Optional[List[torch.Tensor]]:
        """Average ``grads`` (this bucket's gradients, in bucket order) over all ranks, in place.
        Must only enqueue work ordered after the caller's current stream; must not block the host.
        May return replacement gradient tensors (views of the reduced bucket, gradient_as_bucket_view)
        which the Reducer then installs as ``param.grad``; ``None`` means the given tensors were updated."""
        raise NotImplementedError

    def finalize(self) -> None:
        """Make the caller's current stream wait for every bucket enqueued this iteration."""
        raise NotImplementedError

    def sync_order(self, order: List[int]) -> List[int]:
        """Return rank 0's ``order`` on every rank (host-side, once)."""
        raise NotImplementedError

    def broadcast_tensors(self, tensors: Sequence[torch.Tensor], root: int = 0) -> None:
        """Overwrite ``tensors`` on every rank with root's values."""
        raise NotImplementedError

    def close(self) -> None:
        pass


class P2PEngine(BucketEngine):
    """The product engine: buckets live in the IPC-shared arena of a ``P2PComm`` and are reduced by the
    sm_100a kernels of libb200ddp.so on a dedicated high-priority side stream."""

    #: default size of the arena's broadcast staging region when the caller does not size it
    DEFAULT_STAGING = 1 << 20

    def __init__(self, device: torch.device, arena_bytes: int, comm: Optional[P2PComm] = None,
                 staging_bytes: Optional[int] = None) -> None:
        if device.type != "cuda":
            raise RuntimeError("b200ddp reduces gradients on B200 GPUs only; module parameters must be on a "
                               f"CUDA device (got {device}) — there is no CPU path")
        self.device = device
        #: an engine that built its own communicator may replace it with a larger one when a re-bucketing
        #: needs more arena than the construction-time plan did (tests inject loop-back communicators: fixed)
        self._owns_comm = comm is None
        self.comm = comm if comm is not None else P2PComm.from_process_group(arena_bytes, device.index)
        self.rank, self.world_size = self.comm.rank, self.comm.world
        # arena = [broadcast staging | buckets]: parameter/buffer broadcasts never touch bucket storage, so a
        # bucket kernel without a trailing barrier (fused path) cannot be overtaken by a staging write
        self.staging_bytes = (min(self.DEFAULT_STAGING, self.comm.arena_bytes // 4) if staging_bytes is None
                              else int(staging_bytes)) // 4096 * 4096
        self.scale = 1.0 / self.world_size
        with torch.cuda.device(device):
            self.stream = torch.cuda.Stream(device=device, priority=-1)
        self._stream_ptr = self.stream.cuda_stream
        self.max_ctas = int(os.environ.get("B200DDP_CTAS", "0"))
        self._tables: Dict[int, TensorTable] = {}
        self._order_seq = 0
        self._fused_max = min(P2PComm.fused_max_bytes(), int(os.environ.get("B200DDP_FUSED_MAX", str(1 << 30))))
        self._max_table = P2PComm.max_table_tensors()
        self._keepalive: List[torch.Tensor] = []
        self._bcast_plans: Dict[tuple, tuple] = {}
        #: gradient_as_bucket_view (reducer.hpp:238,328-334): after the allreduce the parameters' ``.grad``
        #: become views of the reduced bucket in the arena instead of being copied back (no unflatten pass)
        self.grad_views = False
        self._view_cache: Dict[int, List[torch.Tensor]] = {}
        #: optimizer epilogue (opt-in, DistributedDataParallel.fuse_optimizer): the learning rate the kernels that
        #: write the averaged gradient back also apply to the parameter (p <- p - lr g), or None; and the reduced
        #: parameters in Reducer order (set by the Reducer)
        self.sgd_lr: Optional[float] = None
        self.params: Sequence[torch.Tensor] = ()

    # ---- planning -----------------------------------------------------------------------------------
    def plan(self, buckets: List[BucketSpec]) -> None:
        layout = BucketLayout(start=self.staging_bytes)
        multi = self.world_size > 1
        for b in buckets:
            b.fused = b.nbytes <= self._fused_max and len(b.param_indices) <= self._max_table
            b.algo = _lib.ALGO_ONESHOT if b.fused else self.comm.pick_algo(b.nbytes)
            if not multi:
                b.algo = _lib.ALGO_AUTO
            if b.fused:
                layout.place(b, region_bytes=self.comm.fused_region_bytes(b.nbytes))
            else:  # two regions for the out-of-place one-shot kernel
                layout.place(b, out_of_place=multi and b.algo == _lib.ALGO_ONESHOT)
        if layout.total_bytes > self.comm.arena_bytes:
            self._grow_arena(layout.total_bytes)
        self._tables.clear()
        self._view_cache.clear()
        if multi:
            # (Re-)laying out the arena is collective and fenced on both sides.  Barrier 1: every rank has
            # finished every kernel of the OLD layout (the fused kernels have no trailing barrier, so a slow
            # peer may still be reading this rank's staging slot when this rank's kernel has already ended) —
            # only then may the bytes be reused.  Then the fused path's tagged slots are zeroed.  Barrier 2:
            # every rank's slots are cleared before any peer pushes into them.
            self.comm.stream_wait(self._stream_ptr, torch.cuda.current_stream(self.device).cuda_stream)
            self.comm.barrier(self._stream_ptr)
            for b in buckets:
                if b.fused:
                    self.comm.clear(b.arena_off, self.comm.fused_region_bytes(b.nbytes), self._stream_ptr)
            self.comm.barrier(self._stream_ptr)

    def _grow_arena(self, need: int) -> None:
        """Collective (every rank computes the same layout, so every rank takes this branch together): replace
        the communicator by one whose arena holds ``need`` bytes.  The arena is sized from the construction-time
        plan (arena_bytes_for) — 1x the gradients, like torch's bucket storage — and only a re-bucketing that
        needs more pays this once."""
        if not self._owns_comm:
            raise RuntimeError(f"bucket layout needs {need} bytes but the arena has {self.comm.arena_bytes}")
        import torch.distributed as dist

        torch.cuda.synchronize(self.device)
        if self.world_size > 1 and dist.is_initialized():
            dist.barrier()  # no peer still reads the old arena
        self.comm.destroy()
        self.comm = P2PComm.from_process_group(need + need // 16 + (2 << 20), self.device.index)
        self._bcast_plans.clear()

    def owns(self, t: Optional[torch.Tensor]) -> bool:
        """Is ``t`` a view of this rank's arena (a gradient served as a bucket view)?"""
        if t is None or not t.is_cuda or self.comm is None:
            return False
        lo = self.comm.arena_ptr
        return lo <= t.data_ptr() < lo + self.comm.arena_bytes

    def _table(self, bucket: BucketSpec, tensors: Sequence[torch.Tensor]) -> TensorTable:
        ptrs = [t.data_ptr() for t in tensors]
        cached = self._tables.get(bucket.index)
        want_params = self.sgd_lr is not None
        if cached is not None and cached.key == tuple(ptrs) and (cached.params is not None) == want_params:
            return cached
        params = [self.params[i].data_ptr() for i in bucket.param_indices] if want_params else None
        table = TensorTable(ptrs, bucket.offsets_el, bucket.lengths_el, params)
        self._tables[bucket.index] = table
        return table

    # ---- hot path -----------------------------------------------------------------------------------
    def _bucket_views(self, bucket: BucketSpec, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        views = self._view_cache.get(bucket.index)
        if views is None:
            es = bucket.elem_size
            views = [self.comm.view(bucket.out_off + off * es, n * es, g.dtype).view(g.shape)
                     for off, n, g in zip(bucket.offsets_el, bucket.lengths_el, like)]
            self._view_cache[bucket.index] = views
        return views

    def reduce_bucket(self, bucket: BucketSpec, grads: List[torch.Tensor]) -> Optional[List[torch.Tensor]]:
        comm, s = self.comm, self._stream_ptr
        work = grads
        if not all(g.is_contiguous() for g in grads):
            # rare layout (e.g. channels_last grads): reduce dense copies, copy back on the side stream
            work = [g if g.is_contiguous() else g.contiguous() for g in grads]
        comm.stream_wait(s, torch.cuda.current_stream(self.device).cuda_stream)
        table = self._table(bucket, work)
        if self.grad_views and not bucket.fused:
            # flatten(+scale) into the arena (an in-place scale when .grad already is the view), reduce in
            # place, and hand the arena views back instead of copying out.  Small (fused) buckets keep their
            # own storage: their single launch already writes the result straight into the gradients.
            comm.flatten_scale(table, bucket.dtype_code, bucket.arena_off, self.scale, s)
            if self.world_size > 1:
                comm.allreduce(bucket.arena_off, bucket.out_off, bucket.nbytes, bucket.dtype_code, bucket.algo,
                               self.max_ctas, s)
            self._keepalive.extend(work)  # the flatten kernel on the side stream still reads them
            return self._bucket_views(bucket, grads)
        lr = self.sgd_lr  # optimizer epilogue (parameters are contiguous: checked by fuse_optimizer)
        if bucket.fused:
            if lr is not None:
                comm.fused_bucket_sgd(table, bucket.dtype_code, bucket.arena_off, bucket.nbytes, self.scale, lr, s)
            else:
                comm.fused_bucket(table, bucket.dtype_code, bucket.arena_off, bucket.nbytes, self.scale, s)
        else:
            comm.flatten_scale(table, bucket.dtype_code, bucket.arena_off, self.scale, s)
            if self.world_size > 1:
                comm.allreduce(bucket.arena_off, bucket.out_off, bucket.nbytes, bucket.dtype_code, bucket.algo,
                               self.max_ctas, s)
            if lr is not None:
                comm.unflatten_sgd(table, bucket.dtype_code, bucket.out_off, lr, s)
            else:
                comm.unflatten(table, bucket.dtype_code, bucket.out_off, s)
        if work is not grads:
            with torch.cuda.stream(self.stream):
                for g, w in zip(grads, work):
                    if w is not g:
                        g.copy_(w)
            self._keepalive.extend(work)
        return None

    def finalize(self) -> None:
        self.comm.stream_wait(torch.cuda.current_stream(self.device).cuda_stream, self._stream_ptr)
        self._keepalive.clear()
        self.comm.poll_error()

    def check_health(self, sync: bool = False) -> None:
        """Raise if any kernel of this rank timed out on a peer.  ``sync=True`` first waits for the device (call
        it before writing a checkpoint: a poisoned step must never be saved); without it this is one host read
        (called at the start of every forward, so the exception surfaces at most one step late)."""
        if self.comm is None:
            return
        if sync:
            torch.cuda.synchronize(self.device)
        self.comm.poll_error()

    # ---- control plane (host side, rare) -------------------------------------------------------------
    def sync_order(self, order: List[int]) -> List[int]:
        if self.world_size == 1 or self.comm.uid is None:
            return list(order)  # single rank, or an in-process loop-back world (tests)
        import torch.distributed as dist

        store = dist.distributed_c10d._get_default_store()
        key = self._order_key()
        if self.rank == 0:
            store.set(key, ",".join(map(str, order)))
            return list(order)
        raw = bytes(store.get(key)).decode()
        return [int(x) for x in raw.split(",")] if raw else []

    def _order_key(self) -> str:
        # identical on every rank: communicators are created collectively (comm.uid) and every rank
        # asks for the order the same number of times
        self._order_seq += 1
        return f"b200ddp/order/c{self.comm.uid}/n{self._order_seq}"

    def _broadcast_plan(self, tensors: Sequence[torch.Tensor]):
        """Chunk ``tensors`` into staging-sized multi-tensor tables of raw 4- or 2-byte elements (bit exact
        for any dtype).  Cached per set of storage addresses: the per-forward buffer sync of a model reuses
        the same plan every step."""
        key = tuple((t.data_ptr(), t.numel(), t.dtype) for t in tensors)
        plan = self._bcast_plans.get(key)
        if plan is not None:
            return plan
        odd = [t for t in tensors if not t.is_contiguous() or (t.numel() * t.element_size()) % 2]
        odd_ids = {id(t) for t in odd}
        groups = {4: [], 2: []}
        for t in tensors:
            if id(t) in odd_ids:
                continue
            nbytes = t.numel() * t.element_size()
            groups[4 if nbytes % 4 == 0 and t.data_ptr() % 4 == 0 else 2].append(t)
        cap = self.staging_bytes
        chunks = []
        for unit, ts in groups.items():
            code = _lib.F32 if unit == 4 else _lib.BF16
            i = 0
            while i < len(ts):
                chunk, offs, lens, off = [], [], [], 0
                while i < len(ts) and len(chunk) < self._max_table:
                    n_el = ts[i].numel() * ts[i].element_size() // unit
                    padded = (n_el * unit + 15) // 16 * 16
                    if padded > cap:
                        raise RuntimeError(f"a {padded}-byte tensor does not fit the {cap}-byte broadcast staging "
                                           "region of the arena")
                    if chunk and off * unit + padded > cap:
                        break
                    chunk.append(ts[i]); offs.append(off); lens.append(n_el)
                    off += padded // unit
                    i += 1
                chunks.append((code, TensorTable([t.data_ptr() for t in chunk], offs, lens), max(off * unit, 16)))
        plan = (chunks, odd)
        if len(self._bcast_plans) > 8:
            self._bcast_plans.clear()
        self._bcast_plans[key] = plan
        return plan

    def broadcast_tensors(self, tensors: Sequence[torch.Tensor], root: int = 0) -> None:
        """Root's tensors -> every rank through the arena's staging region by peer stores.  Tensors whose byte
        size is odd, or that are not contiguous, are packed into the staging region with a torch copy (as raw
        bytes / as a dense tensor) and travel through the same broadcast kernel."""
        if self.world_size == 1 or not tensors:
            return
        comm, s = self.comm, self._stream_ptr
        chunks, odd = self._broadcast_plan(tensors)
        cur = torch.cuda.current_stream(self.device).cuda_stream
        comm.stream_wait(s, cur)
        for code, table, nbytes in chunks:
            if self.rank == root:
                comm.flatten_scale(table, code, 0, 1.0, s)
            comm.broadcast(0, nbytes, root, s, self.max_ctas)
            if self.rank != root:
                comm.unflatten(table, code, 0, s)
        if odd:
            with torch.cuda.stream(self.stream):
                for t in odd:
                    self._broadcast_packed(t, root)
        comm.stream_wait(cur, s)

    def _broadcast_packed(self, t: torch.Tensor, root: int) -> None:
        """One tensor of any layout / byte size: root packs it densely into the staging region (chunk by chunk),
        the broadcast kernel moves the bytes, the other ranks copy them out.  Runs on the side stream."""
        comm, s = self.comm, self._stream_ptr
        dense = t if t.is_contiguous() else t.contiguous()
        raw = dense.reshape(-1).view(torch.uint8)
        cap = self.staging_bytes
        for lo in range(0, raw.numel(), cap):
            n = min(cap, raw.numel() - lo)
            stage = comm.view(0, n, torch.uint8)
            if self.rank == root:
                stage.copy_(raw[lo:lo + n])
            comm.broadcast(0, (n + 15) // 16 * 16, root, s, self.max_ctas)
            if self.rank != root:
                raw[lo:lo + n].copy_(stage)
        if dense is not t and self.rank != root:
            t.copy_(dense)

    def close(self) -> None:
        if self.comm is not None:
            torch.cuda.synchronize(self.device)
            self.comm.destroy()
            self.comm = None


class Reducer:
    """Hooks, pending counts, in-order bucket launch, finalize, one-time rebuild."""

    def __init__(
        self,
        params: Sequence[torch.nn.Parameter],
        engine: BucketEngine,
        bucket_bytes_cap: int = DEFAULT_BUCKET_BYTES_CAP,
        first_bucket_bytes_cap: int = DEFAULT_FIRST_BUCKET_BYTES,
    ) -> None:
        self.params = list(params)
        if not self.params:
            raise RuntimeError("DistributedDataParallel is not needed when a module doesn't have any parameter "
                               "that requires a gradient.")
        self.engine = engine
        self.bucket_bytes_cap = int(bucket_bytes_cap)
        self.first_bucket_bytes_cap = int(first_bucket_bytes_cap)
        self._sizes = [p.numel() * p.element_size() for p in self.params]
        self._keys = [(p.dtype, p.device) for p in self.params]
        # construction-time buckets: parameters in definition order, bucket list reversed so that bucket 0
        # holds the last-defined parameters, whose gradients arrive first
        # (torch:nn/parallel/distributed.py:1240-1257)
        indices, _ = compute_bucket_assignment_by_size(
            self._sizes, self._keys, [self.first_bucket_bytes_cap, self.bucket_bytes_cap])
        self.buckets: List[BucketSpec] = []
        self._install(list(reversed(indices)))
        if hasattr(engine, "sgd_lr"):
            engine.params = self.params
        #: optimizer epilogue (DistributedDataParallel.fuse_optimizer): a callable returning the current learning
        #: rate, or None
        self.sgd_lr_source = None
        self._sgd_armed = False
        self._sgd_mark = -1  # the iteration count when the epilogue was last armed
        self.has_rebuilt = False
        self._rebuild_pending = False
        self.armed = False
        self.require_sync = True
        self._callback_queued = False
        self._ready_order: List[int] = []
        self._unfinished = False
        self.iterations = 0
        # The per-parameter hook loop runs in C++ on the autograd thread (csrc/native_reducer.cpp), as torch's own
        # Reducer does, whenever the extension can be loaded; B200DDP_NATIVE=0 (or an engine other than the
        # product engine: the CPU host-logic tests) selects the same state machine in Python below.
        self._native = None
        self._native_ctx = None
        self._hooks = []
        mod = _native.load_native() if (_native.wanted() and isinstance(engine, P2PEngine)) else None
        if mod is not None:
            self._native_ctx = engine.comm.ctx_address
            self._native = mod.NativeReducer(self._native_ctx, engine.comm.function_table(), engine.world_size,
                                             engine._stream_ptr, engine.max_ctas, self.params)
            self._native.configure(self._native_buckets())
            self._hook_stream = self._current_stream()
        else:
            self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    # ---- bucket bookkeeping -------------------------------------------------------------------------
    def _install(self, bucket_indices: List[List[int]]) -> None:
        specs = []
        for bi, idxs in enumerate(bucket_indices):
            p0 = self.params[idxs[0]]
            specs.append(layout_bucket(bi, idxs, [self.params[i].numel() for i in idxs], dtype_code(p0.dtype),
                                       p0.element_size()))
        self.engine.plan(specs)
        self.buckets = specs
        self._param_bucket = [0] * len(self.params)
        for b in specs:
            for i in b.param_indices:
                self._param_bucket[i] = b.index
        if getattr(self, "_native", None) is not None:
            if self.engine.comm.ctx_address != self._native_ctx:  # the engine regrew its arena
                self._native_ctx = self.engine.comm.ctx_address
                self._native.rebind(self._native_ctx)
            self._native.configure(self._native_buckets())
        self._pending = [0] * len(specs)
        self._bucket_ready = [False] * len(specs)
        self._param_seen = [False] * len(self.params)
        self._next_bucket = 0

    def _native_buckets(self):
        eng = self.engine
        out = []
        for b in self.buckets:
            views = []
            if eng.grad_views and not b.fused:
                views = eng._bucket_views(b, [self.params[i] for i in b.param_indices])
            out.append((b.param_indices, b.offsets_el, b.lengths_el, b.dtype_code, b.nbytes, b.arena_off, b.out_off,
                        b.algo, b.fused, views))
        return out

    def _current_stream(self) -> int:
        p = self.params[0]
        return torch.cuda.current_stream(p.device).cuda_stream if p.is_cuda else 0

    def bind_stream(self) -> None:
        """Called at the start of every forward.  The C++ loop keeps the parameters' AccumulateGrad nodes alive
        (its post hooks live on them), and such a node runs on the stream that was current when it was created.
        When the forward's stream changes — eager steps on the default stream, then a CUDA-graph capture on
        torch's capture stream — the nodes are re-created under the new stream, so the whole backward stays on one
        stream (a node left on the legacy default stream would invalidate the capture)."""
        if self._native is None:
            return
        cur = self._current_stream()
        if cur != self._hook_stream:
            self._native.rehook()
            self._hook_stream = cur

    def abort_iteration(self) -> None:
        """Forget an iteration that was armed but will never finish (an aborted CUDA-graph capture): the next
        forward starts clean.  Nothing was executed on the device, so no rank is left waiting."""
        self.armed = False
        self._unfinished = False
        self._callback_queued = False
        if self._native is not None:
            self._native.abort()

    def disarm(self) -> None:
        """The coming backward must not reduce (no_sync, or a forward without grad)."""
        self.armed = False
        if self._native is not None:
            self._native.disarm()

    def bucket_sizes(self) -> List[int]:
        return [b.payload_bytes for b in self.buckets]

    # ---- per-iteration protocol ----------------------------------------------------------------------
    def _arm_sgd(self) -> None:
        """Decide, per iteration, whether the bucket kernels also apply the optimizer step: only when every
        gradient starts from None (then the bucket holds THE gradient of this step; with gradients carried over
        from earlier backwards — accumulation — the optimizer must see the accumulated value and steps itself)."""
        lr = None
        if self.sgd_lr_source is not None and self.require_sync and all(p.grad is None for p in self.params):
            lr = float(self.sgd_lr_source())
        self._sgd_armed = lr is not None
        self.engine.sgd_lr = lr
        if self._native is not None:
            self._native.set_sgd(lr if lr is not None else 0.0)

    def consume_fused_update(self) -> bool:
        """True exactly once after a backward whose kernels applied the optimizer step (the patched
        ``optimizer.step`` then has nothing left to do)."""
        it = int(self._native.iterations()) if self._native is not None else self.iterations
        if self._sgd_armed and it > self._sgd_mark:  # the armed backward has finished (finalize counted it)
            self._sgd_armed = False
            return True
        return False

    def prepare_for_backward(self) -> None:
        """Arm the reducer for the backward of the forward that just ran
        (Reducer::prepare_for_backward, called from DDP._post_forward, distributed.py:1648)."""
        if self.sgd_lr_source is not None or self._sgd_armed:
            self._sgd_mark = int(self._native.iterations()) if self._native is not None else self.iterations
            self._arm_sgd()
        if self._native is not None:
            if not self.has_rebuilt and self._native.iterations() >= 1:
                # the one-time rebuild, driven from here like torch drives it from the next forward
                self._ready_order = list(self._native.take_ready_order())
                self._rebuild_buckets()
            self.iterations = int(self._native.iterations())
            if self.require_sync:
                self._native.prepare()
            else:
                self._native.disarm()
            return
        if self._rebuild_pending:
            self._rebuild_pending = False
            self._rebuild_buckets()
        if self._unfinished:
            raise RuntimeError(
                "Expected to have finished reduction in the prior iteration before starting a new one. "
                "Some parameters did not receive a gradient; b200ddp (like the reference's "
                "DDP(model, device_ids=[gpu_id]), find_unused_parameters=False) requires every parameter "
                "to take part in the loss.")
        for b in self.buckets:
            self._pending[b.index] = len(b.param_indices)
            self._bucket_ready[b.index] = False
        for i in range(len(self._param_seen)):
            self._param_seen[i] = False
        self._next_bucket = 0
        self._callback_queued = False
        self.armed = self.require_sync

    def _make_hook(self, index: int):
        def hook(param: torch.Tensor) -> None:
            self._autograd_hook(index)
        return hook

    def _autograd_hook(self, index: int) -> None:
        if not self.armed:
            return
        if not self._callback_queued:
            self._callback_queued = True
            self._unfinished = True
            Variable._execution_engine.queue_callback(self._finalize_backward)
        if self._param_seen[index]:
            raise RuntimeError(f"parameter {index} was marked ready twice in one backward; b200ddp does not "
                               "support re-entrant backward or shared use of a parameter across graphs")
        self._param_seen[index] = True
        if not self.has_rebuilt:
            self._ready_order.append(index)
        bi = self._param_bucket[index]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._bucket_ready[bi] = True
            n = len(self.buckets)
            while self._next_bucket < n and self._bucket_ready[self._next_bucket]:
                b = self.buckets[self._next_bucket]
                new = self.engine.reduce_bucket(b, [self.params[i].grad for i in b.param_indices])
                if new is not None:  # gradient_as_bucket_view: .grad becomes a view of the reduced bucket
                    for i, g in zip(b.param_indices, new):
                        self.params[i].grad = g
                self._next_bucket += 1

    def _finalize_backward(self) -> None:
        if not self.armed:
            return
        self.armed = False
        if self._next_bucket != len(self.buckets):
            missing = [i for i, seen in enumerate(self._param_seen) if not seen]
            # leave _unfinished set: the next forward raises, like DDP does
            self.engine.finalize()
            raise RuntimeError(f"backward finished but {len(missing)} parameter(s) never produced a gradient "
                               f"(indices {missing[:8]}...); find_unused_parameters is not supported")
        self.engine.finalize()
        self._unfinished = False
        self.iterations += 1
        if not self.has_rebuilt:
            # the one-time rebuild happens at the next forward (as torch drives it, distributed.py:1584): with
            # gradient_as_bucket_view the optimizer still reads .grad views of the CURRENT layout after this callback
            self._rebuild_pending = True

    def _rebuild_buckets(self) -> None:
        """One-time re-bucketing in gradient-ready order (rank 0's order wins on every rank)."""
        order = self.engine.sync_order(self._ready_order)
        self._ready_order = []
        self.has_rebuilt = True
        if sorted(order) != list(range(len(self.params))):
            return  # incomplete order (should not happen once finalize succeeded): keep current buckets
        owns = getattr(self.engine, "owns", None)
        if owns is not None and getattr(self.engine, "grad_views", False):
            # gradients still alive as views of the OLD layout (no zero_grad since the last step: gradient
            # accumulation) move to private storage before the arena is laid out anew
            for p in self.params:
                if owns(p.grad):
                    p.grad = p.grad.clone()
        sizes = [self._sizes[i] for i in order]
        keys = [self._keys[i] for i in order]
        indices, _ = compute_bucket_assignment_by_size(
            sizes, keys, [self.first_bucket_bytes_cap, self.bucket_bytes_cap], tensor_indices=order)
        self._install(indices)

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._native is not None:
            self._native.remove_hooks()
            self._native = None


def arena_bytes_for(params: Sequence[torch.nn.Parameter], world_size: int = 1,
                    bucket_bytes_cap: int = DEFAULT_BUCKET_BYTES_CAP,
                    first_bucket_bytes_cap: int = DEFAULT_FIRST_BUCKET_BYTES) -> int:
    """Bucket storage for the CONSTRUCTION-TIME plan of ``params`` (the same bucketing and region rules
    ``Reducer`` / ``P2PEngine.plan`` apply), plus 1/16 + 2 MiB of slack for the one-time re-bucketing: 1x the
    gradients for in-place buckets (like torch's bucket storage), two slots only for buckets on the fused
    single-launch path (<= 256 KiB) and the LL slot sets only for buckets <= 32 KiB.  A re-bucketing that needs
    more makes the engine replace its arena once (P2PEngine._grow_arena)."""
    lib = _lib.load()
    fused_max = min(int(lib.b200ddp_fused_max_bytes()), int(os.environ.get("B200DDP_FUSED_MAX", str(1 << 30))))
    max_table = int(lib.b200ddp_max_table_tensors())
    out_of_place = os.environ.get("B200DDP_ALGO", "") == str(_lib.ALGO_ONESHOT) or \
        int(os.environ.get("B200DDP_ONESHOT_MAX", "0") or 0) > 0
    params = list(params)
    sizes = [p.numel() * p.element_size() for p in params]
    keys = [(p.dtype, p.device) for p in params]
    indices, _ = compute_bucket_assignment_by_size(sizes, keys, [int(first_bucket_bytes_cap), int(bucket_bytes_cap)])
    align = BucketLayout.REGION_ALIGN
    total = 0
    for idxs in indices:
        es = params[idxs[0]].element_size()
        nbytes = max(16, sum((params[i].numel() * es + 15) // 16 * 16 for i in idxs))
        if nbytes <= fused_max and len(idxs) <= max_table:
            region = max(nbytes, int(lib.b200ddp_fused_region_bytes_for(int(world_size), nbytes)))
        else:
            region = (nbytes + align - 1) // align * align * (2 if (out_of_place and world_size > 1) else 1)
        total += (region + align - 1) // align * align
    return total + total // 16 + (2 << 20)


def staging_bytes_for(tensors: Sequence[torch.Tensor]) -> int:
    """Broadcast staging region: holds the largest single tensor, at least 1 MiB, at most 64 MiB unless one
    tensor is larger (broadcasts are chunked through it)."""
    largest = max((t.numel() * t.element_size() for t in tensors), default=0)
    total = sum((t.numel() * t.element_size() + 15) // 16 * 16 for t in tensors)
    want = max(1 << 20, min(total, 64 << 20), largest + 4096)
    return (want + 4095) // 4096 * 4096
