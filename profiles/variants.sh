#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
SSTB200_CREATE_TIMING=1 python profiles/e2e_breakdown.py --pin 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e.log 2>gpurun_out/bench_e2e.err; tail -c 700 gpurun_out/bench_e2e.log; tail -3 gpurun_out/bench_e2e.err
python bench.py --workload phold --x 1024 --y 1024 --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_phold.log 2>&1; tail -c 700 gpurun_out/bench_phold.log
