#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e.log 2>gpurun_out/bench_e2e.err; tail -c 600 gpurun_out/bench_e2e.log; tail -3 gpurun_out/bench_e2e.err
