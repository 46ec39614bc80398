#!/bin/bash
# one `ncu --set full` capture of the clock-population window kernel (4096^2) + DRAM traffic of a few launches
ARGS="--workload clockpop --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_clockpop.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:dispatch_kernel -s 4 -c 1 -f -o gpurun_out/${1:-prof_clockpop} python bench.py $ARGS > gpurun_out/ncu_clockpop.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dispatch_kernel -s 4 -c 3 --csv --log-file gpurun_out/traffic_clockpop.csv python bench.py $ARGS > /dev/null 2>&1
tail -3 gpurun_out/traffic_clockpop.csv | cut -d, -f12-15
