#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_checkpoint.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu.log
