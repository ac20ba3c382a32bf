#!/bin/bash
# round-2 GPU call 5 (1 GPU): ncu evidence — copy kernels on the real ResNet-50 bucket, the capture driven through
# multigpu_profile.Trainer(profile=True), launch list of the replayed toy step, bench N=1
mkdir -p gpurun_out
timeout 300 python tools/prof_host_path.py > gpurun_out/r24_host_path.txt 2>&1; echo "host path rc=$?"; grep -v Warning gpurun_out/r24_host_path.txt | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:multi_tensor_copy -s 10 -c 6 -f -o gpurun_out/r24_prof_copy python tools/prof_kernels.py > gpurun_out/r24_ncu_copy.log 2>&1; echo "ncu copy rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:multi_tensor_copy -c 20 -f -o gpurun_out/r24_profile_script python tools/prof_profile_script.py > gpurun_out/r24_ncu_profile_script.log 2>&1; echo "ncu profile-script rc=$?"; tail -3 gpurun_out/r24_ncu_profile_script.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r24_launches_bench.csv python bench.py --steps 20 --warmup 5 --blocks 2 --no-sweep --no-extra --no-nccl --no-cpu-baseline > gpurun_out/r24_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r24_bench_n1.json 2> gpurun_out/r24_bench_n1.err; echo "bench N=1 rc=$?"
timeout 600 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r24_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r24_pytest.log
