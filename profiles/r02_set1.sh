#!/bin/bash
# round-2 measurement set on one B200: parity tests, default bench line (both arms), launch list, DRAM traffic and one
# full ncu capture of window_parallel_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_reference.json
ARGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4096.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
bash profiles/ncu_traffic.sh > gpurun_out/traffic.out 2>&1; tail -12 gpurun_out/traffic.out
bash profiles/ncu_full.sh prof_r02_wpk > gpurun_out/full.out 2>&1; tail -2 gpurun_out/full.out
