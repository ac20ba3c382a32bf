"""Shared body of the ``Trainer`` classes the surface scripts expose.

One optimisation step is the reference's ``_run_batch`` (ref:multigpu.py:38-43): zero_grad -> forward
-> loss -> backward -> step.  The gradient allreduce fires inside ``loss.backward()`` through the hooks
``b200ddp.DistributedDataParallel`` installed — nothing here knows about it, exactly as in the reference.

What this module adds around that step (none of it changes what a step computes):

* **CUDA-graph replay of the step** — once the step has run ``GRAPH_AFTER`` times eagerly on batches of one
  shape, it is captured (forward, loss, backward with the bucket kernels our Reducer enqueues on its side stream,
  optimizer step) and later batches of that shape replay it with one launch.  This is the default of the script
  Trainers (``B200DDP_GRAPH=auto``); ``B200DDP_GRAPH=0`` keeps every step eager, ``=1`` captures on the first
  call.  A batch of another shape (the last, partial batch of a shard) runs eagerly; a model that cannot be
  captured (data-dependent control flow, a non-capturable optimizer) falls back to eager for good.  Capture works
  because the allreduce kernels keep their barrier generation in device memory (csrc/kernels.cu).
* **a pipelined host path** — batches that arrive as HOST tensors (the reference's ``DataLoader(pin_memory=True)``,
  ref:multigpu.py:72-79) go through two alternating graphs, each on its OWN stream, that START with the
  host->device copy of their own pinned staging buffer and END with the device->host copy of the loss.  Inside the
  graphs the training step of graph k waits for an external event graph k-1 records after ITS step, so the step
  chain is strictly ordered while the H2D copy of batch i+1 and the loss D2H of step i-1 run under the compute of
  step i (measured: 28.7 us per step with copies and step serialised in one stream, 18 us for the step alone).  The
  loss of step i is read when step i+1 has been queued (``PendingLoss``).  ``B200DDP_HOST_PIPELINE=serial`` keeps
  both graphs on one stream.
* **a device-resident shard**  — ``B200DDP_RESIDENT=1``: the rank's ``DistributedSampler`` shard of
  ``MyTrainDataset`` is moved to HBM once and the captured step gathers its batch by a device-side cursor, so a
  step is exactly one graph launch and no per-batch ``source.to(gpu_id)`` (ref:multigpu.py:48-50).
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, Dict, Optional, Tuple

import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler


def toy_train_objs() -> Tuple[Dataset, torch.nn.Module, torch.optim.Optimizer]:
    """The objects every toy script trains (ref:multigpu.py:65-69, ref:multigpu_torchrun.py:71-75): 2048
    ``MyTrainDataset`` rows, ``Linear(20, 1)``, SGD with lr 1e-3 built on the UN-wrapped model."""
    from .utils import MyTrainDataset

    net = torch.nn.Linear(MyTrainDataset.FEATURES, MyTrainDataset.TARGETS)
    return MyTrainDataset(2048), net, torch.optim.SGD(net.parameters(), lr=1e-3)


def sharded_loader(dataset: Dataset, batch_size: int) -> DataLoader:
    """Per-rank loader (ref:multigpu.py:72-79): pinned memory, no loader-level shuffle, rows chosen by a
    default ``DistributedSampler`` (which shuffles with seed 0 and is never re-seeded)."""
    return DataLoader(dataset, batch_size=batch_size, sampler=DistributedSampler(dataset), shuffle=False,
                      pin_memory=True)


def _truthy(value: Optional[str]) -> bool:
    return value is not None and value.lower() not in ("", "0", "false", "no", "off")


class PendingLoss:
    """The loss of a replayed step whose device->host copy is part of the replay: reading it waits for that step
    only (an event recorded right after the replay), not for the device."""

    __slots__ = ("_host", "_event")

    def __init__(self, host, event: torch.cuda.Event) -> None:
        self._host, self._event = host, event  # host: a 1-element numpy view of the pinned loss word

    def item(self) -> float:
        self._event.synchronize()
        return float(self._host[0])

    __float__ = item


class _Slot:
    """One captured step together with the static buffers it reads and writes."""

    __slots__ = ("graph", "source", "targets", "loss", "host_pack", "dev_pack", "host_source", "host_targets",
                 "host_loss", "host_loss_np", "done", "src_ptr", "tgt_ptr", "src_bytes", "tgt_bytes", "stream", "events",
                 "aliases", "raw")

    def __init__(self) -> None:
        self.graph = None
        self.done = None
        self.stream = None  # host pipeline, overlap mode: the stream this slot's graph is captured and replayed on
        self.raw = None  # (launch function, cudaGraphExec_t, cudaStream_t, cudaEvent_t) when launched through the C ABI



def _signature(source: torch.Tensor, targets: torch.Tensor):
    return (tuple(source.shape), source.dtype, tuple(targets.shape), targets.dtype)


class TrainerCore:
    #: loss used by ``_run_batch``.  The reference's choice (ref:multigpu.py:41) — note SURVEY §9 Q1: with
    #: the toy Linear(20,1) head this loss, and therefore every gradient, is identically zero.
    loss_fn: Callable = staticmethod(torch.nn.CrossEntropyLoss())

    #: eager steps of one batch shape before the step is captured in ``auto`` mode (they are real training steps;
    #: they also carry the Reducer's one-time bucket rebuild, which must precede capture)
    GRAPH_AFTER = 3

    model: torch.nn.Module
    optimizer: torch.optim.Optimizer
    train_data: DataLoader
    gpu_id: int

    graph_kernel_launches: int = 0  # b200ddp kernels captured per replayed step

    # ---- construction --------------------------------------------------------------------------------
    def _adopt(self, wrapped: torch.nn.Module, train_data, optimizer: torch.optim.Optimizer, gpu_id: int,
               save_every: int) -> None:
        """Common constructor tail of the script Trainers."""
        self.gpu_id, self.save_every = gpu_id, save_every
        self.train_data, self.optimizer = train_data, optimizer
        self.model = wrapped
        if _truthy(os.environ.get("B200DDP_FUSE_SGD")) and hasattr(wrapped, "fuse_optimizer"):
            wrapped.fuse_optimizer(optimizer)  # opt-in: the bucket kernels also apply the SGD step (ddp.py)

    def _state(self) -> Dict:
        st = self.__dict__.get("_graph_state")
        if st is None:
            st = self.__dict__["_graph_state"] = {
                "mode": os.environ.get("B200DDP_GRAPH", "auto").lower(), "sig": None, "eager_run": 0,
                "slots": {}, "failed": None, "host_step": 0, "resident": None}
            if st["mode"] in ("1", "on", "true", "yes"):
                st["mode"] = "on"
            elif st["mode"] in ("0", "off", "false", "no"):
                st["mode"] = "off"
            elif st["mode"] != "auto":
                raise ValueError(f"B200DDP_GRAPH must be auto, 0 or 1 (got {st['mode']!r})")
        return st

    @property
    def _graph(self):
        """The captured step for device-resident batches, if any (kept for callers of round 1's API)."""
        slot = self._state()["slots"].get("dev")
        return None if slot is None else slot.graph

    @_graph.setter
    def _graph(self, value) -> None:
        st = self._state()
        if value is None:  # drop every captured step: the next calls run eagerly (and may re-capture)
            st["slots"].clear()
            st["sig"], st["eager_run"], st["mode"] = None, 0, "off"
        else:
            slot = st["slots"].setdefault("dev", _Slot())
            slot.graph = value

    # ---- the step ------------------------------------------------------------------------------------
    def _device(self) -> torch.device:
        gpu = getattr(self, "gpu_id", None)
        if gpu is None:
            return next(self.model.parameters()).device
        return torch.device("cuda", gpu) if isinstance(gpu, int) else torch.device(gpu)

    def _eager_step(self, source: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        self.optimizer.zero_grad()
        loss = self.loss_fn(self.model(source), targets)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _eager_any(self, source: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if not source.is_cuda:  # ref:multigpu.py:48-50
            dev = self._device()
            source, targets = source.to(dev, non_blocking=True), targets.to(dev, non_blocking=True)
        return self._eager_step(source, targets)

    def _run_batch(self, source: torch.Tensor, targets: torch.Tensor):
        """One optimisation step on ``(source, targets)`` — device tensors as in the reference, or the loader's
        host tensors.  Returns the loss (a 0-dim tensor, or a ``PendingLoss`` on the pipelined host path)."""
        st = self._state()
        if st["mode"] == "off" or st["failed"] is not None:
            return self._eager_any(source, targets)
        if st["slots"] and st.get("guard") != self._capture_guard():
            # the optimizer's hyper-parameters (a learning-rate schedule) or the module's train/eval mode changed:
            # a captured step has the old values baked in — drop it and capture again after fresh eager steps
            self._drop_captures()
        sig = _signature(source, targets)
        if st["sig"] != sig:
            if st["slots"]:
                # a batch shaped unlike the captured one (the partial batch that ends a shard): eager this once
                self._join_host_pipeline()
                return self._eager_any(source, targets)
            st["sig"], st["eager_run"] = sig, 0
        kind = "dev" if source.is_cuda else "host"
        if kind != "host":
            self._join_host_pipeline()
        if kind not in st["slots"]:
            # eager steps first: GRAPH_AFTER of them in auto mode, and never fewer than two (the Reducer's one-time
            # bucket rebuild happens at the second forward and must precede capture)
            need = self.GRAPH_AFTER if st["mode"] == "auto" else 2
            if st["eager_run"] < need:
                st["eager_run"] += 1
                if "rng_free" not in st and torch.cuda.is_available():
                    # does the step consume random numbers (dropout)?  torch's replay() advances the captured
                    # generators' offsets; only a step that uses none may be launched through the raw handle
                    dev = self._device()
                    before = torch.cuda.get_rng_state(dev)
                    out = self._eager_any(source, targets)
                    st["rng_free"] = bool(torch.equal(before, torch.cuda.get_rng_state(dev)))
                    return out
                return self._eager_any(source, targets)
            if not self._try_capture(kind, source, targets):
                return self._eager_any(source, targets)
        if kind == "dev":
            slot = st["slots"]["dev"]
            slot.source.copy_(source, non_blocking=True)
            slot.targets.copy_(targets, non_blocking=True)
            slot.graph.replay()
            return slot.loss
        slots = st["slots"]["host"]
        slot = slots[st["host_step"] & 1]
        st["host_step"] += 1
        if slot.stream is not None:
            if not st.get("in_host_pipeline"):
                # entering the two-stream pipeline: both streams start after everything queued so far
                cur = torch.cuda.current_stream(slot.dev_pack.device)
                for s in slots:
                    s.stream.wait_stream(cur)
                st["in_host_pipeline"] = True
            return self._host_step_overlapped(slot, source, targets)
        slot.done.synchronize()  # the replay that last read this staging buffer (two steps ago) has finished
        if source.is_contiguous() and targets.is_contiguous():
            # plain memcpy of the loader's (pinned) batch into the staging buffer: ~1 us instead of two tensor copies
            ctypes.memmove(slot.src_ptr, source.data_ptr(), slot.src_bytes)
            ctypes.memmove(slot.tgt_ptr, targets.data_ptr(), slot.tgt_bytes)
        else:
            slot.host_source.copy_(source)
            slot.host_targets.copy_(targets)
        slot.graph.replay()
        slot.done.record()
        return PendingLoss(slot.host_loss_np, slot.done)

    def _capture_guard(self):
        """What a captured step has baked in besides shapes: the optimizer's scalar hyper-parameters (they are
        kernel arguments of the captured optimizer step) and the module's training flag."""
        groups = getattr(self.optimizer, "param_groups", ())
        return (getattr(self.model, "training", True),) + tuple(
            (g.get("lr"), g.get("momentum"), g.get("weight_decay"), g.get("dampening"), g.get("nesterov")) for g in groups)

    def _drop_captures(self) -> None:
        self._join_host_pipeline()
        dev = self._device()
        torch.cuda.synchronize(dev)
        st = self._state()
        st["slots"].clear()
        st["resident"] = None
        st["sig"], st["eager_run"], st["host_step"] = None, 0, 0

    def _host_step_overlapped(self, slot: "_Slot", source: torch.Tensor, targets: torch.Tensor) -> PendingLoss:
        slot.done.synchronize()  # this slot's previous replay (two steps ago) has released the staging buffer
        if source.is_contiguous() and targets.is_contiguous():
            ctypes.memmove(slot.src_ptr, source.data_ptr(), slot.src_bytes)
            ctypes.memmove(slot.tgt_ptr, targets.data_ptr(), slot.tgt_bytes)
        else:
            slot.host_source.copy_(source)
            slot.host_targets.copy_(targets)
        raw = slot.raw
        if raw is not None:
            # cudaGraphLaunch + cudaEventRecord on the slot's stream in one C call (b200ddp_launch_graph): ~3 us of
            # host time instead of ~18 us for `with stream: replay(); record()`, which made this path host-bound
            rc = raw[0](raw[1], raw[2], raw[3])
            if rc:
                from . import _lib

                _lib.check(rc, None, "launch_graph")
        else:
            with torch.cuda.stream(slot.stream):
                slot.graph.replay()
                slot.done.record()
        return PendingLoss(slot.host_loss_np, slot.done)

    def _join_host_pipeline(self) -> None:
        """Leaving the two-stream host pipeline (an eager step, another replay kind, a checkpoint): the current
        stream continues after both pipeline streams."""
        st = self._state()
        if st.get("in_host_pipeline"):
            slots = st["slots"].get("host") or []
            for s in slots:
                if s.stream is not None:
                    torch.cuda.current_stream(s.dev_pack.device).wait_stream(s.stream)
            st["in_host_pipeline"] = False

    # ---- capture -------------------------------------------------------------------------------------
    def _capture_step(self, prologue: Callable[[], Tuple[torch.Tensor, torch.Tensor]],
                      epilogue: Optional[Callable[[torch.Tensor], None]] = None, stream=None):
        """Capture prologue -> forward -> loss -> backward (with our bucket kernels) -> optimizer step -> epilogue.
        All ranks must call this collectively: the captured allreduce kernels synchronise across ranks."""
        dev = self._device()
        torch.cuda.synchronize(dev)
        comm = getattr(getattr(self.model, "engine", None), "comm", None)
        n0 = comm.launch_count if comm is not None else 0
        self.optimizer.zero_grad(set_to_none=True)  # gradients are (re)created inside the graph's pool
        graph = torch.cuda.CUDAGraph()
        with (torch.cuda.graph(graph) if stream is None else torch.cuda.graph(graph, stream=stream)):
            source, targets = prologue()
            loss = self.loss_fn(self.model(source), targets)
            loss.backward()
            self.optimizer.step()
            out = loss.detach()
            if epilogue is not None:
                epilogue(out)
        self.graph_kernel_launches = (comm.launch_count - n0) if comm is not None else 0
        return graph, out

    def _try_capture(self, kind: str, source: torch.Tensor, targets: torch.Tensor) -> bool:
        st = self._state()
        try:
            if kind == "dev":
                st["slots"]["dev"] = self._capture_dev(source, targets)
            else:
                how = os.environ.get("B200DDP_HOST_PIPELINE", "overlap").lower()
                if how not in ("overlap", "serial"):
                    raise ValueError(f"B200DDP_HOST_PIPELINE must be overlap or serial (got {how!r})")
                if how == "overlap":
                    dev = self._device()
                    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
                    # external events become event-record / event-wait NODES of the graphs: the step of graph k
                    # waits for the step of graph k-1 wherever the two replays run
                    stepped = [torch.cuda.Event(external=True) for _ in range(2)]
                    cur = torch.cuda.current_stream(dev)
                    for ev in stepped:
                        ev.record(cur)
                    st["slots"]["host"] = [self._capture_host(source, targets, stream=streams[k],
                                                              wait_for=stepped[k ^ 1], signal=stepped[k])
                                           for k in range(2)]
                else:
                    st["slots"]["host"] = [self._capture_host(source, targets) for _ in range(2)]
            st["guard"] = self._capture_guard()
            if _truthy(os.environ.get("B200DDP_VERBOSE")):
                print(f"[GPU{getattr(self, 'gpu_id', '?')}] training step captured into a CUDA graph (kind={kind}, "
                      f"b200ddp kernels per replay={self.graph_kernel_launches})", flush=True)
            return True
        except Exception as exc:  # the model / optimizer cannot be captured: this Trainer stays eager
            st["failed"] = exc
            st["slots"].pop(kind, None)
            reducer = getattr(self.model, "reducer", None)
            if reducer is not None and hasattr(reducer, "abort_iteration"):
                reducer.abort_iteration()
            try:
                torch.cuda.synchronize(self._device())
            except Exception:
                pass
            if _truthy(os.environ.get("B200DDP_GRAPH_STRICT")):
                raise
            print(f"[GPU{getattr(self, 'gpu_id', '?')}] CUDA-graph capture of the training step failed "
                  f"({type(exc).__name__}: {exc}); continuing eagerly")
            return False

    def _capture_dev(self, source: torch.Tensor, targets: torch.Tensor) -> _Slot:
        slot = _Slot()
        slot.source, slot.targets = source.clone(), targets.clone()
        slot.graph, slot.loss = self._capture_step(lambda: (slot.source, slot.targets))
        return slot

    def _capture_host(self, source: torch.Tensor, targets: torch.Tensor, stream=None, wait_for=None,
                      signal=None) -> _Slot:
        """A step that begins with the H2D copy of its own pinned staging buffer (source and targets packed into
        one transfer) and ends with the D2H copy of the loss.  Overlap mode (``stream`` given): captured on its own
        stream; between the H2D copy and the step the graph waits for ``wait_for`` (the other graph's step), after
        the optimizer step it records ``signal`` — only then comes the loss D2H, off the step chain."""
        dev = self._device()
        slot = _Slot()
        sb = source.numel() * source.element_size()
        tb = targets.numel() * targets.element_size()
        t_off = (sb + 255) // 256 * 256
        slot.host_pack = torch.empty(t_off + max(tb, 1), dtype=torch.uint8).pin_memory()
        slot.dev_pack = torch.empty(t_off + max(tb, 1), dtype=torch.uint8, device=dev)
        slot.host_source = slot.host_pack[:sb].view(source.dtype).view(source.shape)
        slot.host_targets = slot.host_pack[t_off:t_off + tb].view(targets.dtype).view(targets.shape)
        slot.source = slot.dev_pack[:sb].view(source.dtype).view(source.shape)
        slot.targets = slot.dev_pack[t_off:t_off + tb].view(targets.dtype).view(targets.shape)
        slot.host_source.copy_(source)
        slot.host_targets.copy_(targets)
        slot.host_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
        slot.host_loss_np = slot.host_loss.numpy()
        slot.src_ptr, slot.tgt_ptr = slot.host_source.data_ptr(), slot.host_targets.data_ptr()
        slot.src_bytes, slot.tgt_bytes = sb, tb
        slot.done = torch.cuda.Event()
        # Small batches (the toy rows: 2.7 KB) cross PCIe inside copy KERNELS that read / write the pinned buffers
        # through their device mapping (pinned memory is device-addressable under unified addressing): a
        # copy-engine memcpy node costs ~7 us of graph time per transfer at this size, a kernel node ~2 us.  Large
        # batches (ResNet-50: 19 MB) keep the copy engine.  B200DDP_STAGE=memcpy|kernel forces one.
        how = os.environ.get("B200DDP_STAGE", "auto").lower()
        # overlap mode keeps the copy engine for every size: the copies are off the step chain there, and a
        # copy-engine transfer runs beside the step's kernels instead of taking an SM
        mapped = how == "kernel" or (how != "memcpy" and stream is None and slot.host_pack.numel() <= (64 << 10))
        slot.stream = stream
        slot.events = (wait_for, signal)  # the graph's event nodes refer to them: they live as long as the graph
        src_alias = loss_alias = None
        if mapped:
            from .comm import _CudaArray

            src_alias = torch.as_tensor(_CudaArray(slot.host_pack.data_ptr(), slot.host_pack.numel()), device=dev)
            loss_alias = torch.as_tensor(_CudaArray(slot.host_loss.data_ptr(), 4), device=dev).view(torch.float32)
        slot.aliases = (src_alias, loss_alias)

        def prologue():
            if mapped:
                slot.dev_pack.copy_(src_alias)  # H2D: a kernel reading the pinned staging buffer over PCIe
            else:
                slot.dev_pack.copy_(slot.host_pack, non_blocking=True)  # H2D: copy-engine memcpy node
            if wait_for is not None:
                torch.cuda.current_stream(dev).wait_event(wait_for)  # event-wait node: the previous step is done
            return slot.source, slot.targets

        def epilogue(loss: torch.Tensor) -> None:
            if signal is not None:
                signal.record(torch.cuda.current_stream(dev))  # event-record node: this step is done
            if mapped:
                loss_alias.copy_(loss.float().reshape(1))  # D2H: a 4-byte store into pinned host memory
            else:
                slot.host_loss.copy_(loss.float().reshape(1), non_blocking=True)

        slot.graph, slot.loss = self._capture_step(prologue, epilogue, stream=stream)
        if stream is not None:
            slot.done.record(stream)
            if (self._state().get("rng_free") and hasattr(slot.graph, "raw_cuda_graph_exec")
                    and os.environ.get("B200DDP_RAW_LAUNCH", "1") not in ("0", "false", "no")):
                from . import _lib

                try:
                    slot.raw = (_lib.load().b200ddp_launch_graph, int(slot.graph.raw_cuda_graph_exec()),
                                int(stream.cuda_stream), int(slot.done.cuda_event))
                except (TypeError, ValueError, RuntimeError):
                    slot.raw = None  # this torch does not hand out the handle as an integer: torch's replay() it is
        else:
            slot.done.record()
        return slot

    def enable_cuda_graph(self, source: torch.Tensor, targets: torch.Tensor, warmup: int = 3) -> None:
        """Capture the step NOW for batches shaped like (source, targets) (device tensors).  The ``warmup`` eager
        steps on that batch are real training steps (at least two run unless the Trainer has already stepped:
        they carry the Reducer's one-time bucket rebuild).  Collective across ranks."""
        st = self._state()
        if "dev" in st["slots"]:
            return
        self._join_host_pipeline()
        dev = source.device
        src, tgt = source.clone(), targets.clone()
        need = max(2, warmup) if st["eager_run"] < 2 else warmup
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(need):
                self._eager_step(src, tgt)
        torch.cuda.current_stream(dev).wait_stream(side)
        st["mode"] = "on" if st["mode"] == "off" else st["mode"]
        st["failed"] = None
        st["sig"] = _signature(source, targets)
        st["slots"]["dev"] = self._capture_dev(source, targets)
        st["guard"] = self._capture_guard()

    # ---- device-resident shard -----------------------------------------------------------
    def enable_resident_shard(self, loader=None) -> bool:
        """Keep this rank's sampler shard in HBM (``utils.DeviceShardLoader``) and capture a step that gathers
        its batch from it by a device-side cursor.  ``loader``: a ``DeviceShardLoader``, or None to build one from
        ``self.train_data`` (a ``DataLoader`` over a ``MyTrainDataset`` with a ``DistributedSampler``).  Returns
        False (and changes nothing) when the data cannot be made resident.  Collective across ranks."""
        from .utils import DeviceShardLoader

        st = self._state()
        if st["resident"] is not None:
            return True
        self._join_host_pipeline()
        if loader is None:
            data = self.train_data
            ds = getattr(data, "dataset", None)
            if ds is None or not hasattr(ds, "device_shard") or data.batch_size is None:
                return False
            loader = DeviceShardLoader(ds, data.batch_size, self._device(), sampler=getattr(data, "sampler", None))
        bsz = loader.batch_size
        rows = loader.x.shape[0]
        full = rows // bsz * bsz
        if full == 0:
            return False
        dev = loader.x.device
        if st["eager_run"] < 2:  # the Reducer's one-time rebuild must precede capture
            for _ in range(2):
                self._eager_step(loader.x[:bsz], loader.y[:bsz])
            st["eager_run"] += 2
        res = {"loader": loader, "bsz": bsz, "full": full, "steps": full // bsz,
               "cursor": torch.zeros((), dtype=torch.int64, device=dev)}
        x, y = loader.x, loader.y
        packed = (x.dim() == 2 and y.dim() == 2 and x.dtype == y.dtype and x.element_size() == 4)
        if packed:
            # rows of [source | targets] gathered by ONE launch of our gather kernel (b200ddp_gather_rows), which
            # also advances the cursor: the replayed step starts with a single extra node
            from . import _lib

            lib = _lib.load()
            fx = x.shape[1]
            res["xy"] = torch.cat([x[:full], y[:full]], dim=1).contiguous()
            res["bx"] = torch.empty(bsz, fx, dtype=x.dtype, device=dev)
            res["by"] = torch.empty(bsz, y.shape[1], dtype=y.dtype, device=dev)

            def prologue():
                rc = lib.b200ddp_gather_rows(res["xy"].data_ptr(), res["bx"].data_ptr(), res["by"].data_ptr(),
                                             res["xy"].shape[1] * 4, fx * 4, bsz, full, res["cursor"].data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(rc, None, "gather_rows")
                return res["bx"], res["by"]

            epilogue = None
        else:
            res["lane"] = torch.arange(bsz, dtype=torch.int64, device=dev)

            def prologue():
                idx = res["lane"] + res["cursor"]
                return x.index_select(0, idx), y.index_select(0, idx)

            def epilogue(_loss: torch.Tensor) -> None:
                res["cursor"].add_(bsz).remainder_(full)

        try:
            res["graph"], res["loss"] = self._capture_step(prologue, epilogue)
        except Exception as exc:
            st["failed"] = exc
            reducer = getattr(self.model, "reducer", None)
            if reducer is not None and hasattr(reducer, "abort_iteration"):
                reducer.abort_iteration()
            return False
        if packed:
            self.graph_kernel_launches += 1  # gather_rows_kernel is ours too (it has no context to count it)
        res["launches_per_replay"] = self.graph_kernel_launches
        if not st["slots"]:
            st["guard"] = self._capture_guard()
        res["cursor"].zero_()
        st["resident"] = res
        return True

    def _run_resident_step(self) -> torch.Tensor:
        """One step on the next resident batch: a single graph launch."""
        self._join_host_pipeline()
        res = self._state()["resident"]
        res["graph"].replay()
        return res["loss"]

    def _run_resident_epoch(self) -> None:
        self._join_host_pipeline()
        res = self._state()["resident"]
        res["cursor"].zero_()
        for _ in range(res["steps"]):
            res["graph"].replay()
        loader = res["loader"]
        if loader.x.shape[0] > res["full"]:  # the partial batch that ends the shard (the loader keeps it)
            self._eager_step(loader.x[res["full"]:], loader.y[res["full"]:])

    # ---- epoch loop ------------------------------------------------------------------------------------
    def _run_epoch(self, epoch: int) -> None:
        steps = len(self.train_data)
        first = next(iter(self.train_data))[0]
        print(f"[GPU{self.gpu_id}] Epoch {epoch} | Batchsize: {len(first)} | Steps: {steps}")
        st = self._state()
        if _truthy(os.environ.get("B200DDP_RESIDENT")) and st["mode"] != "off":
            if st["resident"] is not None and st.get("guard") != self._capture_guard():
                self._drop_captures()  # e.g. a scheduler changed the learning rate between epochs: capture anew
            if self.enable_resident_shard():
                self._run_resident_epoch()
                return
        for source, targets in self.train_data:
            # host batches go in as they come: the eager path moves them (ref:multigpu.py:48-50), the replayed
            # path stages them itself
            self._run_batch(source, targets)

    def _check_health(self) -> None:
        """Before a checkpoint: wait for the device and raise if any bucket kernel timed out on a peer."""
        self._join_host_pipeline()
        check = getattr(self.model, "check_health", None)
        if check is not None:
            check(True)
