#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_cli.py tests/test_gpu_checkpoint.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_multi.log
