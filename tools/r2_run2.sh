#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_capture.py > gpurun_out/r21_debug_capture.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r21_debug_capture.txt
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout=600 > gpurun_out/r21_pytest_ddp.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r21_pytest_ddp.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r21_smoke.log 2>&1; echo "smoke rc=$?"
