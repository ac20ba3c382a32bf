"""``P2PComm`` — Python owner of one ``b200ddp_ctx`` (include/b200ddp.h): the IPC-shared bucket arena of
this rank, the mappings of every peer's arena, and thin methods over the C ABI's hot-path calls.

Bootstrap replaces what ``init_process_group(backend="nccl")`` sets up for the allreduce
(ref:multigpu.py:12-20): each rank exports a 128-byte handle, the handles travel through the c10d
store (host-side rendezvous only), and ``b200ddp_open_peers`` maps the peers over NVLink.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from typing import List, Optional, Sequence

import torch

from . import _lib

_DTYPE_CODES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
_comm_seq = itertools.count()


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPE_CODES[dtype]
    except KeyError:
        raise TypeError(f"b200ddp reduces float32/bfloat16/float16 gradients, got {dtype}") from None


def _mem_mode_from_env() -> int:
    name = os.environ.get("B200DDP_MEM", "auto").lower()
    modes = {"auto": _lib.MEM_AUTO, "ipc": _lib.MEM_IPC, "vmm": _lib.MEM_VMM}
    if name not in modes:
        raise ValueError(f"B200DDP_MEM must be one of {sorted(modes)}, got {name!r}")
    return modes[name]


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier so torch can view library-owned device memory."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 2,
        }


class TensorTable:
    """ctypes arrays (pointers, element offsets, element counts) for one multi-tensor call."""

    __slots__ = ("n", "ptrs", "offs", "lens", "key", "params")

    def __init__(self, data_ptrs: Sequence[int], offsets_el: Sequence[int], lengths_el: Sequence[int],
                 param_ptrs: Optional[Sequence[int]] = None) -> None:
        n = len(data_ptrs)
        self.n = n
        self.ptrs = (C.c_void_p * n)(*data_ptrs)
        self.offs = (C.c_uint64 * n)(*offsets_el)
        self.lens = (C.c_uint64 * n)(*lengths_el)
        self.key = tuple(data_ptrs)
        #: storage of the parameter each tensor is the gradient of (optimizer epilogue, opt-in)
        self.params = None if param_ptrs is None else (C.c_void_p * n)(*param_ptrs)


class P2PComm:
    """One rank's context.  ``stream`` arguments are raw ``cudaStream_t`` integers (``Stream.cuda_stream``)."""

    def __init__(self, rank: int, world: int, device: int, arena_bytes: int, mem_mode: Optional[int] = None) -> None:
        self._lib = _lib.load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        ctx = C.c_void_p()
        mode = _mem_mode_from_env() if mem_mode is None else mem_mode
        rc = self._lib.b200ddp_init(self.rank, self.world, self.device, int(arena_bytes), mode, C.byref(ctx))
        _lib.check(rc, None, "init")
        self._ctx = ctx
        self.arena_bytes = int(self._lib.b200ddp_arena_bytes(ctx))
        ptr = C.c_void_p()
        _lib.check(self._lib.b200ddp_arena_ptr(ctx, C.byref(ptr)), ctx, "arena_ptr")
        self.arena_ptr = int(ptr.value)
        self._arena_u8: Optional[torch.Tensor] = None
        self.uid: Optional[int] = None  # set when built collectively over a process group
        timeout = os.environ.get("B200DDP_TIMEOUT_MS")
        if timeout:
            self.set_timeout_ms(int(timeout))

    # ---- bootstrap ----------------------------------------------------------------------------------
    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(_lib.HANDLE_BYTES)
        _lib.check(self._lib.b200ddp_export_handle(self._ctx, buf, _lib.HANDLE_BYTES), self._ctx, "export_handle")
        return buf.raw

    def open_peers(self, handles: Sequence[bytes]) -> None:
        blob = b"".join(handles)
        if len(blob) != _lib.HANDLE_BYTES * self.world:
            raise ValueError("need one handle per rank")
        _lib.check(self._lib.b200ddp_open_peers(self._ctx, blob, self.world), self._ctx, "open_peers")

    @classmethod
    def from_process_group(cls, arena_bytes: int, device: Optional[int] = None, group=None) -> "P2PComm":
        """Collective constructor for the one-process-per-GPU case: exchanges handles through the c10d
        store of the default process group (TCP rendezvous; no collective on the data path)."""
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("call init_process_group (ddp_setup) before building a b200ddp communicator")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()
        comm = cls(rank, world, device, arena_bytes)
        comm.uid = -1
        if world > 1:
            store = dist.distributed_c10d._get_default_store()
            # the sequence number is identical on all ranks because communicators are built collectively
            comm.uid = next(_comm_seq)
            prefix = f"b200ddp/comm{comm.uid}"
            store.set(f"{prefix}/h{rank}", comm.export_handle())
            handles = [bytes(store.get(f"{prefix}/h{r}")) for r in range(world)]
            comm.open_peers(handles)
        return comm

    @classmethod
    def loopback_world(cls, world: int, device: int, arena_bytes: int, mem_mode: int = _lib.MEM_IPC) -> List["P2PComm"]:
        """``world`` ranks inside THIS process, all on ``device`` (1-GPU parity tests): peers are plain
        pointers, the kernels and the barrier protocol are exactly the multi-GPU ones."""
        comms = [cls(r, world, device, arena_bytes, mem_mode) for r in range(world)]
        handles = [c.export_handle() for c in comms]
        for c in comms:
            c.open_peers(handles)
        return comms

    # ---- arena access -------------------------------------------------------------------------------
    def arena(self) -> torch.Tensor:
        """The whole arena as a uint8 CUDA tensor (a view of library-owned memory)."""
        if self._arena_u8 is None:
            self._arena_u8 = torch.as_tensor(_CudaArray(self.arena_ptr, self.arena_bytes),
                                             device=torch.device("cuda", self.device))
        return self._arena_u8

    def view(self, off: int, nbytes: int, dtype: torch.dtype) -> torch.Tensor:
        return self.arena()[off:off + nbytes].view(dtype)

    # ---- hot path -----------------------------------------------------------------------------------
    def flatten_scale(self, table: TensorTable, dtype: int, bucket_off: int, scale: float, stream: int) -> None:
        rc = self._lib.b200ddp_flatten_scale(self._ctx, table.ptrs, table.offs, table.lens, table.n, dtype,
                                             bucket_off, scale, stream)
        if rc:
            _lib.check(rc, self._ctx, "flatten_scale")

    def unflatten(self, table: TensorTable, dtype: int, bucket_off: int, stream: int) -> None:
        rc = self._lib.b200ddp_unflatten(self._ctx, table.ptrs, table.offs, table.lens, table.n, dtype, bucket_off,
                                         stream)
        if rc:
            _lib.check(rc, self._ctx, "unflatten")

    def allreduce(self, in_off: int, out_off: int, nbytes: int, dtype: int, algo: int, max_ctas: int,
                  stream: int, scale: float = 1.0) -> None:
        rc = self._lib.b200ddp_allreduce(self._ctx, in_off, out_off, nbytes, dtype, algo, max_ctas, scale, stream)
        if rc:
            _lib.check(rc, self._ctx, "allreduce")

    def fused_bucket(self, table: TensorTable, dtype: int, bucket_off: int, bucket_bytes: int, scale: float,
                     stream: int) -> None:
        rc = self._lib.b200ddp_fused_bucket(self._ctx, table.ptrs, table.offs, table.lens, table.n, dtype,
                                            bucket_off, bucket_bytes, scale, stream)
        if rc:
            _lib.check(rc, self._ctx, "fused_bucket")

    def broadcast(self, off: int, nbytes: int, root: int, stream: int, max_ctas: int = 0) -> None:
        rc = self._lib.b200ddp_broadcast(self._ctx, off, nbytes, root, max_ctas, stream)
        if rc:
            _lib.check(rc, self._ctx, "broadcast")

    def peer_read(self, peer: int, src_off: int, dst_off: int, nbytes: int, stream: int, max_ctas: int = 0) -> None:
        _lib.check(self._lib.b200ddp_peer_read(self._ctx, peer, src_off, dst_off, nbytes, max_ctas, stream),
                   self._ctx, "peer_read")

    def unflatten_sgd(self, table: TensorTable, dtype: int, bucket_off: int, lr: float, stream: int) -> None:
        rc = self._lib.b200ddp_unflatten_sgd(self._ctx, table.ptrs, table.params, table.offs, table.lens, table.n, dtype,
                                             bucket_off, lr, stream)
        if rc:
            _lib.check(rc, self._ctx, "unflatten_sgd")

    def fused_bucket_sgd(self, table: TensorTable, dtype: int, bucket_off: int, bucket_bytes: int, scale: float,
                         lr: float, stream: int) -> None:
        rc = self._lib.b200ddp_fused_bucket_sgd(self._ctx, table.ptrs, table.params, table.offs, table.lens, table.n,
                                                dtype, bucket_off, bucket_bytes, scale, lr, stream)
        if rc:
            _lib.check(rc, self._ctx, "fused_bucket_sgd")

    def fused_region_bytes(self, bucket_bytes: int) -> int:
        return int(self._lib.b200ddp_fused_region_bytes(self._ctx, bucket_bytes))

    def clear(self, off: int, nbytes: int, stream: int) -> None:
        _lib.check(self._lib.b200ddp_clear(self._ctx, off, nbytes, stream), self._ctx, "clear")

    def barrier(self, stream: int) -> None:
        _lib.check(self._lib.b200ddp_barrier(self._ctx, stream), self._ctx, "barrier")

    def stream_wait(self, waiter: int, signaled: int) -> None:
        rc = self._lib.b200ddp_stream_wait(self._ctx, waiter, signaled)
        if rc:
            _lib.check(rc, self._ctx, "stream_wait")

    # ---- diagnostics --------------------------------------------------------------------------------
    def poll_error(self) -> None:
        rc = self._lib.b200ddp_poll_error(self._ctx)
        if rc:
            _lib.check(rc, self._ctx, "kernel")

    def set_timeout_ms(self, ms: int) -> None:
        _lib.check(self._lib.b200ddp_set_timeout_ms(self._ctx, ms), self._ctx, "set_timeout_ms")

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200ddp_launch_count(self._ctx))

    @property
    def has_multicast(self) -> bool:
        return bool(self._lib.b200ddp_has_multicast(self._ctx))

    @property
    def mem_mode(self) -> str:
        return {_lib.MEM_IPC: "ipc", _lib.MEM_VMM: "vmm"}.get(self._lib.b200ddp_mem_mode(self._ctx), "?")

    @property
    def ctx_address(self) -> int:
        return int(self._ctx.value)

    def function_table(self) -> dict:
        """name -> address of every C-ABI entry point (for the native hook loop, which calls the same ABI)."""
        return {name: C.cast(getattr(self._lib, name), C.c_void_p).value for name, _, _ in _lib.SYMBOLS}

    def pick_algo(self, nbytes: int) -> int:
        return int(self._lib.b200ddp_pick_algo(self._ctx, nbytes))

    @staticmethod
    def max_table_tensors() -> int:
        return int(_lib.load().b200ddp_max_table_tensors())

    @staticmethod
    def fused_max_bytes() -> int:
        return int(_lib.load().b200ddp_fused_max_bytes())

    def destroy(self) -> None:
        if self._ctx is not None:
            self._arena_u8 = None
            ctx, self._ctx = self._ctx, None
            self._lib.b200ddp_destroy(ctx)

    def __del__(self) -> None:  # best effort; explicit destroy() is the documented path
        try:
            self.destroy()
        except Exception:
            pass
