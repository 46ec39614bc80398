#!/bin/bash
# round-end measurement set on one B200: parity tests, the default bench line (both arms), launch list + DRAM traffic +
# one full ncu capture of dispatch_kernel, PHOLD / clock-population lines and PHOLD traffic
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_reference.json
timeout 600 python bench.py --workload phold --x 1024 --y 1024 --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_phold.json 2>/dev/null; tail -c 400 gpurun_out/bench_phold.json
timeout 600 python bench.py --workload clockpop --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_clockpop.json 2>/dev/null; tail -c 400 gpurun_out/bench_clockpop.json
bash profiles/ncu_traffic.sh > gpurun_out/traffic.out 2>&1
bash profiles/ncu_full.sh prof_v15 > gpurun_out/full.out 2>&1
ARGS="--workload phold --x 1024 --y 1024 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dispatch_kernel -s 4 -c 2 --csv --log-file gpurun_out/traffic_phold.csv python bench.py $ARGS > /dev/null 2>&1
tail -4 gpurun_out/traffic_phold.csv | cut -d, -f13-15
