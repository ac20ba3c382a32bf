"""ctypes binding of ``libb200ddp.so`` — one Python function per entry point of include/b200ddp.h.

There is deliberately no fallback: if the library is missing or a call fails, a ``RuntimeError`` is
raised (the way DDP surfaces reducer errors, torch:nn/parallel/distributed.py:1178-1181).
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_PATH

# constants mirrored from include/b200ddp.h
ABI_VERSION = 2
MAX_WORLD = 8
HANDLE_BYTES = 128
F32, BF16, F16 = 0, 1, 2
MEM_AUTO, MEM_IPC, MEM_VMM = 0, 1, 2
ALGO_AUTO, ALGO_ONESHOT, ALGO_TWOSHOT, ALGO_TWOSHOT_TMA, ALGO_NVLS = 0, 1, 2, 3, 4
ALGO_NAMES = {"auto": 0, "oneshot": 1, "twoshot": 2, "tma": 3, "twoshot_tma": 3, "nvls": 4}
E_TIMEOUT = -4

# every symbol the header declares: (name, restype, argtypes)
_vp, _u64, _i, _f = C.c_void_p, C.c_uint64, C.c_int, C.c_float
_pp = C.POINTER(C.c_void_p)
_pu64 = C.POINTER(C.c_uint64)
SYMBOLS = [
    ("b200ddp_abi_version", _i, []),
    ("b200ddp_init", _i, [_i, _i, _i, _u64, _i, _pp]),
    ("b200ddp_export_handle", _i, [_vp, _vp, _u64]),
    ("b200ddp_open_peers", _i, [_vp, _vp, _i]),
    ("b200ddp_arena_ptr", _i, [_vp, _pp]),
    ("b200ddp_arena_bytes", _u64, [_vp]),
    ("b200ddp_has_multicast", _i, [_vp]),
    ("b200ddp_mem_mode", _i, [_vp]),
    ("b200ddp_destroy", _i, [_vp]),
    ("b200ddp_last_error", C.c_char_p, [_vp]),
    ("b200ddp_flatten_scale", _i, [_vp, _pp, _pu64, _pu64, _i, _i, _u64, _f, _vp]),
    ("b200ddp_unflatten", _i, [_vp, _pp, _pu64, _pu64, _i, _i, _u64, _vp]),
    ("b200ddp_allreduce", _i, [_vp, _u64, _u64, _u64, _i, _i, _i, _f, _vp]),
    ("b200ddp_fused_bucket", _i, [_vp, _pp, _pu64, _pu64, _i, _i, _u64, _u64, _f, _vp]),
    ("b200ddp_unflatten_sgd", _i, [_vp, _pp, _pp, _pu64, _pu64, _i, _i, _u64, _f, _vp]),
    ("b200ddp_fused_bucket_sgd", _i, [_vp, _pp, _pp, _pu64, _pu64, _i, _i, _u64, _u64, _f, _f, _vp]),
    ("b200ddp_max_table_tensors", _i, []),
    ("b200ddp_fused_max_bytes", _u64, []),
    ("b200ddp_fused_region_bytes", _u64, [_vp, _u64]),
    ("b200ddp_fused_region_bytes_for", _u64, [_i, _u64]),
    ("b200ddp_peer_read", _i, [_vp, _i, _u64, _u64, _u64, _i, _vp]),
    ("b200ddp_launch_graph", _i, [_vp, _vp, _vp]),
    ("b200ddp_gather_rows", _i, [_vp, _vp, _vp, _u64, _u64, C.c_uint32, _u64, _vp, _vp]),
    ("b200ddp_clear", _i, [_vp, _u64, _u64, _vp]),
    ("b200ddp_broadcast", _i, [_vp, _u64, _u64, _i, _i, _vp]),
    ("b200ddp_barrier", _i, [_vp, _vp]),
    ("b200ddp_stream_wait", _i, [_vp, _vp, _vp]),
    ("b200ddp_poll_error", _i, [_vp]),
    ("b200ddp_set_timeout_ms", _i, [_vp, C.c_uint32]),
    ("b200ddp_launch_count", _u64, [_vp]),
    ("b200ddp_pick_algo", _i, [_vp, _u64]),
]

_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(b200ddp has no CPU or library fallback for the gradient allreduce path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    got = lib.b200ddp_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libb200ddp.so has ABI version {got}, the Python host expects {ABI_VERSION}")
    _lib = lib
    return lib


def last_error(ctx) -> str:
    msg = load().b200ddp_last_error(ctx)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, ctx=None, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"b200ddp {what} failed (code {rc}): {last_error(ctx)}")
