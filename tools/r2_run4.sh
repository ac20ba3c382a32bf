#!/bin/bash
# round-2 GPU call 4 (8 GPUs): the driver-style bench line at N=8, multi-GPU parity tests at ws=8, configs[2..4]
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/r23_gpus.txt 2>&1
timeout 700 $TR --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r23_bench_n8.json 2> gpurun_out/r23_bench_n8.err; echo "bench N=8 rc=$?"
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout=600 -k "baseline_configs or comm_hook_shim_matches or (raw_allreduce and vmm) or gradient_as_bucket_view_matches" > gpurun_out/r23_pytest_multi.log 2>&1; echo "pytest-multi rc=$?"; grep -E "worst rel|passed|failed" gpurun_out/r23_pytest_multi.log | tail -8
timeout 400 $TR --master-port 29723 bench.py --gpus 8 --workload resnet50 --views --no-sweep > gpurun_out/r23_bench_resnet_views_n8.json 2> gpurun_out/r23_bench_resnet_views_n8.err; echo "resnet views N=8 rc=$?"
timeout 500 $TR --master-port 29725 bench.py --gpus 8 --workload gpt2m --no-sweep > gpurun_out/r23_bench_gpt2m_n8.json 2> gpurun_out/r23_bench_gpt2m_n8.err; echo "gpt2m N=8 rc=$?"
timeout 500 $TR --master-port 29727 tools/cap_sweep.py > gpurun_out/r23_cap_sweep_n8.jsonl 2> gpurun_out/r23_cap_sweep_n8.err; echo "cap sweep rc=$?"
tail -c 400 gpurun_out/r23_bench_n8.err
