#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x --timeout 150 -k periodic > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_multi.log
