"""oracle/ — TEST INFRASTRUCTURE, not product code.

A CPU restatement of the reference's algorithm for the one hot path (the per-step DDP gradient-bucket
allreduce, ref:multigpu.py:36,42) used only as the CHECKER.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import or execute anything here; the
product package (distributed-pytorch_b200/) never does and fails loudly without its CUDA library.

Where the algorithm lives: the reference contributes one line (``DDP(model, device_ids=[gpu_id])``); the
arithmetic is in its only dependency, torch (``torch>=1.11.0``, ref:requirements.txt:1; installed and
pinned here as torch 2.11.0+cu128 with NCCL 2.28.9 / gloo) — c10d::Reducer + ProcessGroup{NCCL,Gloo}.
torch IS present in this container and on the GPU box, so the oracle is pinned two ways:

  * ``ref_ddp.py`` runs the dependency's own implementation (stock
    ``torch.nn.parallel.DistributedDataParallel`` on gloo/CPU, or nccl on the GPU box) on seeded inputs;
  * ``tests/golden/make_golden.py`` ran the reference's own ``Trainer`` (imported from /root/reference,
    only ``DDP``/``init_process_group`` rebound to gloo/CPU) and torch's own
    ``_compute_bucket_assignment_by_size``; the recorded vectors are committed under tests/golden/ and
    every restatement below is checked against them in the CPU test-suite.

The reference itself has no tests, golden vectors or known-answer fixtures for this path (SURVEY §4,
§8c): the pins above are "outputs of the reference stack run here", not reference-authored fixtures.

Restatements:
  bucket_oracle.py      compute_bucket_assignment_by_size (reducer.hpp:592-598; distributed.py:1219-1247)
  allreduce_oracle.py   flatten + 1/ws pre-scale, rank-ordered fp32 SUM, unflatten (reducer.hpp:276,330;
                        default_comm_hooks.hpp:36-51)
  allreduce_ref.c       the same arithmetic in C for full-size checks and the timed CPU baseline
"""
