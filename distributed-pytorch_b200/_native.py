"""Loader of the C++ hook loop (csrc/native_reducer.cpp), built in-tree with torch's C++ extension machinery
(ninja + g++, no nvcc: it only needs the CUDA *stream* types).  It is the default hook loop of ``Reducer`` on the
product engine whenever it can be loaded; ``B200DDP_NATIVE=0`` selects the Python loop of reducer.py.

``build_native()`` (called by ``__graft_entry__.build()``) compiles the extension into ``_native_build/`` and
records the SHA-256 of the source next to the ``.so``; ``load_native()`` imports that ``.so`` directly when the
recorded hash matches the current source — no ninja run, no JIT cache — so the file that was built here is
the file that is loaded on the GPU box (it travels with the tree).
"""
from __future__ import annotations

import hashlib
import importlib.machinery
import importlib.util
import os
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(PKG_DIR, "csrc", "native_reducer.cpp")
BUILD_DIR = os.path.join(PKG_DIR, "_native_build")
NAME = "b200ddp_native"
SO_PATH = os.path.join(BUILD_DIR, NAME + ".so")
STAMP = os.path.join(BUILD_DIR, NAME + ".sha256")

_module = None
_load_failed = None


def _source_hash() -> str:
    import torch

    h = hashlib.sha256()
    with open(SOURCE, "rb") as f:
        h.update(f.read())
    h.update(torch.__version__.encode())  # the extension is ABI-bound to the torch it was compiled against
    return h.hexdigest()


def _prebuilt_ok() -> bool:
    try:
        with open(STAMP) as f:
            return os.path.exists(SO_PATH) and f.read().strip() == _source_hash()
    except OSError:
        return False


def _import_so():
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    loader = importlib.machinery.ExtensionFileLoader(NAME, SO_PATH)
    spec = importlib.util.spec_from_file_location(NAME, SO_PATH, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


def build_native(verbose: bool = False):
    """Compile (if stale) and import the extension; returns the module."""
    global _module
    if _module is not None:
        return _module
    if _prebuilt_ok():
        _module = _import_so()
        return _module
    from torch.utils import cpp_extension

    os.makedirs(BUILD_DIR, exist_ok=True)
    _module = cpp_extension.load(
        name=NAME,
        sources=[SOURCE],
        build_directory=BUILD_DIR,
        extra_cflags=["-O2", "-std=c++17"],
        with_cuda=True,  # include paths + c10_cuda/cudart for the stream types; no .cu source, so no nvcc
        verbose=verbose,
    )
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return _module


def load_native():
    """The extension if it can be had (prebuilt, or buildable here), else ``None`` — the Python hook loop is the
    same state machine and takes over."""
    global _load_failed
    if _module is not None:
        return _module
    if _load_failed is not None:
        return None
    try:
        return build_native()
    except Exception as exc:  # no compiler on this box, ABI mismatch, ...
        _load_failed = exc
        return None


def wanted() -> bool:
    """Default on; ``B200DDP_NATIVE=0`` turns the C++ loop off."""
    return os.environ.get("B200DDP_NATIVE", "1") not in ("", "0", "false", "no")
