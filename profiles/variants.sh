#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_periodic_stats.py -m gpu -q -x --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
