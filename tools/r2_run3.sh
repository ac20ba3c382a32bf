#!/bin/bash
# round-2 GPU call 3 (2 GPUs): real multi-process worlds — the multi-GPU tests, the bench line at N=2, N=1 beside it
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r22_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout=900 -k "multigpu" > gpurun_out/r22_pytest_multi.log 2>&1; echo "pytest-multi rc=$?"; grep -E "worst rel|passed|failed" gpurun_out/r22_pytest_multi.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r22_bench_n2.json 2> gpurun_out/r22_bench_n2.err; echo "bench N=2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r22_bench_n1.json 2> gpurun_out/r22_bench_n1.err; echo "bench N=1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 5 --impl nccl > gpurun_out/r22_nccl_n2.json 2> gpurun_out/r22_nccl_n2.err; echo "nccl N=2 rc=$?"
tail -c 600 gpurun_out/r22_bench_n2.err
