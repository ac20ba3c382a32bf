"""Build / load ``oracle/liboracle.so`` (the C restatement, TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
SRC = os.path.join(HERE, "allreduce_ref.c")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        proc = subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + proc.stdout + proc.stderr)
    return LIB


def load() -> C.CDLL:
    lib = C.CDLL(build())
    pf = C.POINTER(C.c_float)
    ppf = C.POINTER(pf)
    pu = C.POINTER(C.c_uint64)
    lib.oracle_threads.restype = C.c_int
    lib.oracle_flatten_scale.argtypes = [ppf, pu, pu, C.c_int, C.c_float, pf]
    lib.oracle_allreduce_sum.argtypes = [ppf, C.c_int, C.c_uint64, pf]
    lib.oracle_unflatten.argtypes = [pf, pu, pu, C.c_int, ppf]
    lib.oracle_ddp_step.argtypes = [C.POINTER(ppf), pu, pu, C.c_int, C.c_int, C.c_uint64, pf]
    lib.oracle_ddp_step.restype = C.c_uint64
    return lib


if __name__ == "__main__":
    print(build(force=True))
