"""``DistributedDataParallel`` — the drop-in for the one line the reference contributes to the hot path:
``self.model = DDP(model, device_ids=[gpu_id])`` (ref:multigpu.py:36, ref:multigpu_torchrun.py:34,
ref:multinode_torchrun.py:35, ref:multigpu_profile.py:39-42).

It keeps the surface the reference scripts use — ``DDP(module, device_ids=[i])`` / ``DDP(module, [i])``,
``ddp(inputs)``, ``.module``, ``state_dict()`` with ``module.``-prefixed keys (ref:multigpu_profile.py:77),
gradients landing in the ORIGINAL ``Parameter.grad`` objects (the optimizer is built on the inner model
before wrapping, ref:multigpu.py:68,84-86) — and the construction-time semantics of torch's class
(torch:nn/parallel/distributed.py:661-960): parameter-shape verification, rank-0 parameter/buffer
broadcast, per-forward buffer broadcast, ``no_sync()``.  Below that surface everything is ours: the
custom ``Reducer`` and the sm_100a peer-to-peer kernels behind the C ABI.
"""
from __future__ import annotations

import os
from contextlib import contextmanager
from typing import Any, Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

from .bucketing import DEFAULT_BUCKET_BYTES_CAP, DEFAULT_FIRST_BUCKET_BYTES
from .reducer import BucketEngine, P2PEngine, Reducer, arena_bytes_for, staging_bytes_for

_live_ddp: List["DistributedDataParallel"] = []


def _move(obj: Any, device: torch.device) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj if obj.device == device else obj.to(device, non_blocking=True)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_move(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: _move(v, device) for k, v in obj.items()}
    return obj


class DistributedDataParallel(nn.Module):
    def __init__(
        self,
        module: nn.Module,
        device_ids: Optional[Sequence[Any]] = None,
        output_device: Optional[Any] = None,
        dim: int = 0,
        broadcast_buffers: bool = True,
        process_group: Any = None,
        bucket_cap_mb: Optional[float] = None,
        find_unused_parameters: bool = False,
        check_reduction: bool = False,
        gradient_as_bucket_view: bool = False,
        static_graph: bool = False,
        *,
        _engine: Optional[BucketEngine] = None,
    ) -> None:
        super().__init__()
        if find_unused_parameters:
            raise NotImplementedError("find_unused_parameters=True is outside the reference's path "
                                      "(DDP(model, device_ids=[gpu_id])) and is not supported")
        if process_group is not None:
            raise NotImplementedError("b200ddp spans the default process group (one 8-GPU NVSwitch box)")
        if device_ids is not None and len(device_ids) > 1:
            raise ValueError("device_ids takes exactly one device (one process per GPU)")
        self.module = module
        self.dim = dim
        self.broadcast_buffers = broadcast_buffers
        self.static_graph = static_graph
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise RuntimeError("DistributedDataParallel is not needed when a module doesn't have any parameter "
                               "that requires a gradient.")
        devices = {p.device for p in params}
        if len(devices) != 1:
            raise ValueError(f"all parameters must live on one device, got {sorted(map(str, devices))}")
        self.device = next(iter(devices))
        if device_ids is not None:
            want = torch.device(device_ids[0]) if not isinstance(device_ids[0], int) else torch.device("cuda", device_ids[0])
            if _engine is None and want != self.device:
                raise ValueError(f"device_ids={list(device_ids)} but the module's parameters are on {self.device}")
        self.device_ids = list(device_ids) if device_ids is not None else None
        self.output_device = output_device
        cap = DEFAULT_BUCKET_BYTES_CAP if bucket_cap_mb is None else int(bucket_cap_mb * 1024 * 1024)
        env_cap = os.environ.get("B200DDP_BUCKET_MB")
        if env_cap and bucket_cap_mb is None:
            cap = int(float(env_cap) * 1024 * 1024)
        # an explicit cap also applies to the first bucket (torch:nn/parallel/distributed.py:1230-1236,1262-1266)
        first = DEFAULT_FIRST_BUCKET_BYTES if bucket_cap_mb is None and not env_cap else cap
        self.bucket_bytes_cap = cap

        state = self._module_states()
        if _engine is None:
            # arena = broadcast staging (construction-time / per-forward sync) + bucket storage sized from the
            # construction-time bucket plan (1x the gradients + the small fused buckets' second slot)
            import torch.distributed as dist

            world = dist.get_world_size() if dist.is_initialized() else 1
            staging = staging_bytes_for(state)
            self.engine: BucketEngine = P2PEngine(self.device, arena_bytes_for(params, world, cap, first) + staging,
                                                  staging_bytes=staging)
        else:
            self.engine = _engine
        # B200DDP_GRAD_VIEWS=1 turns torch's own flag on without touching the reference's DDP(model, device_ids=[i])
        # call (knobs are environment variables so the script signatures stay identical)
        if os.environ.get("B200DDP_GRAD_VIEWS", "0") not in ("", "0", "false", "no"):
            gradient_as_bucket_view = True
        if gradient_as_bucket_view:
            # .grad of every parameter in a large bucket becomes a view of the reduced bucket in
            # the arena (no unflatten pass); as in torch, such gradients must not be detach_()-ed.  The one-time
            # bucket rebuild happens at the next forward; views still alive then are moved to private storage
            self.engine.grad_views = True
        self.gradient_as_bucket_view = bool(gradient_as_bucket_view)
        self._verify_param_shapes(params)
        # ranks start from different random initialisations (ref:multigpu.py:67); rank 0 wins
        # (torch:nn/parallel/distributed.py:883-890)
        self.engine.broadcast_tensors(state, root=0)
        self.reducer = Reducer(params, self.engine, cap, first)
        self.require_backward_grad_sync = True
        self.require_forward_param_sync = True
        _live_ddp.append(self)

    # ---- construction helpers ------------------------------------------------------------------------
    def _module_states(self) -> List[torch.Tensor]:
        return [p.detach() for p in self.module.parameters()] + [b.detach() for b in self.module.buffers()]

    def _verify_param_shapes(self, params: Iterable[torch.Tensor]) -> None:
        """All ranks must wrap the same architecture (``_verify_param_shape_across_processes``,
        torch:nn/parallel/distributed.py:881).  Host-side, through the c10d store."""
        if self.engine.world_size == 1:
            return
        import hashlib

        import torch.distributed as dist

        if not dist.is_initialized():
            return
        sig = hashlib.sha1(repr([(tuple(p.shape), str(p.dtype)) for p in params]).encode()).hexdigest()
        sigs: List[Any] = [None] * dist.get_world_size()
        store = dist.distributed_c10d._get_default_store()
        uid = getattr(getattr(self.engine, "comm", None), "uid", None)
        prefix = f"b200ddp/shapes/{uid}/{len(_live_ddp)}"
        store.set(f"{prefix}/{dist.get_rank()}", sig)
        for r in range(dist.get_world_size()):
            sigs[r] = bytes(store.get(f"{prefix}/{r}")).decode()
        if any(s != sigs[0] for s in sigs):
            raise RuntimeError("DDP expects the same model across all ranks, but parameter shapes/dtypes differ "
                               f"between ranks (signatures {sigs})")

    # ---- forward -------------------------------------------------------------------------------------
    def fuse_optimizer(self, optimizer: torch.optim.Optimizer) -> bool:
        """Opt-in (second half): let the kernels that write the averaged gradient back also apply
        the optimizer step, ``p <- p - lr * g``, so the reference's ``self.optimizer.step()`` (ref:multigpu.py:43;
        plain ``SGD(lr=1e-3)``, :68) has nothing left to do — one pass over the parameters and one launch less
        per step.  Only plain SGD is fused (no momentum, weight decay, dampening, nesterov or maximize; every
        parameter of the optimizer contiguous and reduced by this DDP; not with ``gradient_as_bucket_view``);
        anything else returns False and changes nothing.

        This moves the update INTO ``backward()``: code that runs between ``backward()`` and ``step()`` (gradient
        clipping, loss scaling, skipping a step) sees gradients whose update has already been applied — which is why
        it is opt-in (``B200DDP_FUSE_SGD=1`` makes the script Trainers call it).  A backward that starts with
        gradients already present (accumulation) or runs under ``no_sync()`` is never fused: the optimizer then
        steps itself.  ``optimizer.step`` is wrapped, the optimizer object stays the caller's."""
        if getattr(self, "_fused_optimizer", None) is not None:
            return self._fused_optimizer is optimizer
        if type(optimizer) is not torch.optim.SGD or self.gradient_as_bucket_view:
            return False
        if not hasattr(self.engine, "sgd_lr"):
            return False
        groups = optimizer.param_groups
        mine = {id(p) for p in self.reducer.params}
        owned = [p for g in groups for p in g["params"]]
        if len(groups) != 1 or {id(p) for p in owned} != mine or not all(p.is_contiguous() for p in owned):
            return False
        g = groups[0]
        if g.get("momentum", 0) or g.get("weight_decay", 0) or g.get("dampening", 0) or g.get("nesterov", False) \
                or g.get("maximize", False) or g.get("differentiable", False):
            return False
        reducer = self.reducer
        reducer.sgd_lr_source = lambda: optimizer.param_groups[0]["lr"]
        plain_step = optimizer.step

        def step(closure=None):
            if closure is None and reducer.consume_fused_update():
                return None  # the bucket kernels of the backward that just finished applied p <- p - lr g
            return plain_step(closure) if closure is not None else plain_step()

        optimizer.step = step
        self._fused_optimizer = optimizer
        return True

    def check_health(self, sync: bool = True) -> None:
        """Raise if a kernel of this rank timed out on a peer (the step's gradients are then NOT the average).
        The script Trainers call this before writing a checkpoint."""
        check = getattr(self.engine, "check_health", None)
        if check is not None:
            check(sync)

    def forward(self, *inputs: Any, **kwargs: Any) -> Any:
        check = getattr(self.engine, "check_health", None)
        if check is not None:
            check(False)  # one host read: a timeout of the previous step surfaces before the next one starts
        self.reducer.bind_stream()
        if self.device_ids is not None:
            inputs, kwargs = _move(inputs, self.device), _move(kwargs, self.device)
        if self.require_forward_param_sync and self.broadcast_buffers:
            # rank 0's buffers before the forward (will_sync_module_buffers / _sync_buffers,
            # torch:nn/parallel/distributed.py:1590-1591,2191-2196,2214-2256).  As in torch, the flag is what
            # the PREVIOUS forward left behind: buffers are synchronised before a forward iff the forward
            # before it was a gradient-synchronising one (:1637-1650)
            buffers = [b.detach() for b in self.module.buffers()]
            if buffers:
                self.engine.broadcast_tensors(buffers, root=0)
        out = self.module(*inputs, **kwargs)
        if torch.is_grad_enabled() and self.require_backward_grad_sync:
            self.require_forward_param_sync = True
            self.reducer.require_sync = True
            self.reducer.prepare_for_backward()
        else:
            self.require_forward_param_sync = False
            if torch.is_grad_enabled():
                self.reducer.require_sync = False
                self.reducer.disarm()
        return out

    @contextmanager
    def no_sync(self):
        """Skip gradient synchronisation inside the context (gradients accumulate locally)."""
        old = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old

    # ---- teardown ------------------------------------------------------------------------------------
    def shutdown(self) -> None:
        """Release the arena (called by ``destroy_process_group``)."""
        if self in _live_ddp:
            _live_ddp.remove(self)
        self.reducer.remove_hooks()
        self.engine.close()


def shutdown_all() -> None:
    for ddp in list(_live_ddp):
        ddp.shutdown()
