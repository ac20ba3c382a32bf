#!/bin/bash
# round-2 GPU call 1 (1 GPU): smoke, the GPU test suite, the driver-style bench line, copy-kernel A/B, smoke under ncu
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r20_gpus.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r20_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s --timeout=600 > gpurun_out/r20_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r20_pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r20_bench.json 2> gpurun_out/r20_bench.err; echo "bench rc=$?"
timeout 300 python tools/ab_copy.py > gpurun_out/r20_ab_copy.txt 2>&1; echo "ab rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r20_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r20_smoke_ncu.log 2>&1; echo "smoke-ncu rc=$?"
tail -3 gpurun_out/r20_smoke_ncu.log
