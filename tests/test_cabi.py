"""The drop-in boundary: libb200ddp.so loads and exports exactly what include/b200ddp.h declares.
No compute calls here (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from b200ddp import _lib
from b200ddp._build import LIB_PATH

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "b200ddp.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ddp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared_functions()
    bound = sorted(name for name, _, _ in _lib.SYMBOLS)
    assert declared == bound, "include/b200ddp.h and _lib.SYMBOLS list different entry points"


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200ddp_\w+)", out))
    assert exported == set(_declared_functions()), "library exports symbols the header does not declare"


def test_header_constants_match_python():
    text = open(HEADER).read()

    def const(name):
        return int(re.search(rf"#define {name} \(?(-?\d+)\)?", text).group(1))

    assert const("B200DDP_ABI_VERSION") == _lib.ABI_VERSION == _lib.load().b200ddp_abi_version()
    assert const("B200DDP_MAX_WORLD") == _lib.MAX_WORLD
    assert const("B200DDP_HANDLE_BYTES") == _lib.HANDLE_BYTES
    assert (const("B200DDP_F32"), const("B200DDP_BF16"), const("B200DDP_F16")) == (_lib.F32, _lib.BF16, _lib.F16)
    assert const("B200DDP_ALGO_NVLS") == _lib.ALGO_NVLS and const("B200DDP_ALGO_TWOSHOT_TMA") == _lib.ALGO_TWOSHOT_TMA
    assert const("B200DDP_ETIMEOUT") == _lib.E_TIMEOUT
    assert _lib.load().b200ddp_max_table_tensors() == 160
    assert _lib.load().b200ddp_fused_max_bytes() == 256 << 10


def test_library_has_no_libcuda_or_torch_dependency():
    out = subprocess.run(["ldd", LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libc10" not in out


def test_library_contains_sm100a_code_with_tma_and_multimem():
    """The hot kernels are real sm_100a SASS: UBLKCP (cp.async.bulk / TMA) and LDGMC (multimem.ld_reduce)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not installed")
    lst = subprocess.run([cuobjdump, "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in lst
    sass = subprocess.run([cuobjdump, "-sass", LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "LDGMC" in sass and "LDG.E.NA.128" in sass and "STG.E.NA.128" in sass


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_init_fails_loudly():
    from b200ddp.comm import P2PComm

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P2PComm(0, 1, 0, 1 << 20)
    # bad arguments are rejected before touching CUDA
    ctx = C.c_void_p()
    assert _lib.load().b200ddp_init(3, 2, 0, 1 << 20, 0, C.byref(ctx)) == -1
    assert "out of range" in _lib.last_error(None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_ddp_refuses_cpu_module():
    from b200ddp import DDP

    with pytest.raises(RuntimeError, match="no CPU path"):
        DDP(torch.nn.Linear(4, 2))


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under distributed-pytorch_b200/ (Python or C/CUDA) may import,
    link or execute it, and the product has no CPU/NCCL implementation of the bucket engine."""
    import glob

    pkg = os.path.join(os.path.dirname(HEADER), "..", "distributed-pytorch_b200")
    files = glob.glob(os.path.join(pkg, "*.py")) + glob.glob(os.path.join(pkg, "csrc", "*"))
    assert files
    for path in files:
        text = open(path).read()
        # (docstrings may MENTION the oracle; what is forbidden is importing, linking or executing it)
        for needle in ("import oracle", "from oracle", "liboracle", "oracle.build", "oracle.ref_ddp", "import_module(\"oracle"):
            assert needle not in text, f"{os.path.basename(path)} references the oracle ({needle!r})"
    reducer = open(os.path.join(pkg, "reducer.py")).read()
    assert "dist.all_reduce" not in reducer and "all_reduce(" not in reducer  # no library collective on the path


def test_profile_script_surface():
    import inspect

    from b200ddp import multigpu_profile as m

    assert list(inspect.signature(m.Trainer.__init__).parameters) == ["self", "model", "train_set", "optimizer",
                                                                      "device", "batch_size", "profile"]
    for name in ("prepare_model_for_DDP", "run_batch", "run_epoch", "save_checkpoint", "_create_profiler"):
        assert callable(getattr(m.Trainer, name))
    assert list(inspect.signature(m.ddp_setup).parameters) == ["rank", "world_size"]
    assert list(inspect.signature(m.main).parameters) == ["rank", "world_size"]
    w = m.CudaProfilerWindow(wait=1, warmup=1, active=5)
    assert (w.first, w.last) == (2, 7)
