#!/bin/bash
# round-2 closing set on one B200: parity tests, the default bench line (both arms), PHOLD / clock-population lines,
# launch list of the default command, DRAM traffic of the three window kernels, full ncu captures of the two new ones,
# RNG streams.  Every ncu run comes after the same command has exited 0 without it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; tail -c 400 gpurun_out/bench_reference.json
timeout 600 python bench.py --workload phold --x 1024 --y 1024 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_phold.json 2>/dev/null; tail -c 300 gpurun_out/bench_phold.json
timeout 600 python bench.py --workload clockpop --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_clockpop.json 2>/dev/null; tail -c 300 gpurun_out/bench_clockpop.json
timeout 300 python profiles/rng_bench.py > gpurun_out/rng_bench.json; python profiles/rng_print.py
# launch list of the default command
A="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-steady"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4096.csv python bench.py $A > gpurun_out/ncu_launches.log 2>&1
# DRAM traffic per launch
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'window_parallel_kernel|maintenance_kernel' -s 14 -c 8 --csv --log-file gpurun_out/traffic_4096.csv python bench.py $A > /dev/null 2>&1
python profiles/traffic_table.py gpurun_out/traffic_4096.csv
P="--workload phold --x 1024 --y 1024 --steps 6 --warmup 10 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:window_ordered -s 12 -c 4 --csv --log-file gpurun_out/traffic_phold.csv python bench.py $P > /dev/null 2>&1
python profiles/traffic_table.py gpurun_out/traffic_phold.csv
C="--workload clockpop --steps 6 --warmup 10 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:window_clock -s 12 -c 4 --csv --log-file gpurun_out/traffic_clockpop.csv python bench.py $C > /dev/null 2>&1
python profiles/traffic_table.py gpurun_out/traffic_clockpop.csv
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_ordered -s 12 -c 1 -f -o gpurun_out/prof_r02_ordered python bench.py $P > gpurun_out/ncu_full_phold.log 2>&1; tail -1 gpurun_out/ncu_full_phold.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_clock -s 12 -c 1 -f -o gpurun_out/prof_r02_clock python bench.py $C > gpurun_out/ncu_full_clock.log 2>&1; tail -1 gpurun_out/ncu_full_clock.log
