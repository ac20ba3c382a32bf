"""Loader of the opt-in C++ hook loop (csrc/native_reducer.cpp), built in-tree with torch's C++ extension
machinery (ninja + g++, no nvcc: it only needs the CUDA *stream* types).  ``B200DDP_NATIVE=1`` selects it in
``Reducer``; without the environment variable the Python hook loop of reducer.py is used."""
from __future__ import annotations

import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(PKG_DIR, "csrc", "native_reducer.cpp")
BUILD_DIR = os.path.join(PKG_DIR, "_native_build")

_module = None


def build_native(verbose: bool = False):
    """Compile (if stale) and import the extension; returns the module."""
    global _module
    if _module is not None:
        return _module
    from torch.utils import cpp_extension

    os.makedirs(BUILD_DIR, exist_ok=True)
    _module = cpp_extension.load(
        name="b200ddp_native",
        sources=[SOURCE],
        build_directory=BUILD_DIR,
        extra_cflags=["-O2", "-std=c++17"],
        with_cuda=True,  # include paths + c10_cuda/cudart for the stream types; no .cu source, so no nvcc
        verbose=verbose,
    )
    return _module


def wanted() -> bool:
    return os.environ.get("B200DDP_NATIVE", "0") not in ("", "0", "false", "no")
