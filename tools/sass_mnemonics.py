"""SASS evidence per kernel: cuobjdump -sass libb200ddp.so, memory / sync mnemonics counted per function."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "distributed-pytorch_b200", "libb200ddp.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
dem = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
KEEP = re.compile(r"^(LDG|STG|LDGMC|UBLKCP|SYNCS|MEMBAR|FENCE|CCTL|LDS\.128|LDC|RED|ATOM|ERRBAR)")
want = [("multi_tensor_copy_kernel<0, false, 4>", "multi_tensor_copy_kernel<0,false,4>"), ("multi_tensor_copy_kernel<0, true, 4>", "multi_tensor_copy_kernel<0,true,4>"),
        ("fused_ll_kernel<0, 8>", None), ("fused_bucket_kernel<0, 8>", None), ("fused_bucket_kernel<0, 1>", None), ("oneshot_kernel<0, 8>", None),
        ("twoshot_kernel<0, 8>", None), ("twoshot_tma_kernel<0, 8>", None), ("nvls_kernel<0, 4>", None), ("nvls_kernel<1, 4>", None),
        ("broadcast_kernel", None), ("barrier_kernel", None), ("local_copy_kernel", None), ("gather_rows_kernel", None)]
print("# SASS evidence (round 2): cuobjdump -sass distributed-pytorch_b200/libb200ddp.so (sm_100a); fp32 instantiations at ws=8 shown.")
print("# UBLKCP.S.G = cp.async.bulk global->shared (TMA bulk copy); SYNCS.* = mbarrier; LDGMC.E.ADD = multimem.ld_reduce;")
print("# STG.E.128.STRONG.SYS (on the multicast address) = multimem.st; LDG/STG.E.NA.128 = 16-byte L1::no_allocate streams;")
print("# LDG/STG.E.STRONG.SYS = acquire/release flag traffic on the peer-mapped signal pads; *.64.STRONG.SYS = LL data+tag words;")
print("# FENCE.VIEW.ASYNC = fence.proxy.async (generic -> async proxy ordering before the first bulk copy of the TMA two-shot);")
print("# LDC with a register index = the copy kernel's CTA-uniform table lookups in the kernel-parameter constant bank.\n")
for needle, label in want:
    for name, ins in zip(dem, funcs.values()):
        if needle in name:
            c = collections.Counter(i for i in ins if KEEP.match(i))
            ldc = c.pop("LDC", 0) + sum(c.pop(k) for k in list(c) if k.startswith("LDC."))
            items = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()))
            print(f"{label or needle.replace(', ', ',')}: {len(ins)} instrs; {items}" + (f", LDC* x{ldc}" if "multi_tensor" in needle else ""))
            break
print(f"\nkernel instantiations in the library: {len(funcs)}")
