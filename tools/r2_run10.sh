#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r33_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r33_pytest.log
