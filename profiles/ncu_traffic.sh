#!/bin/bash
# DRAM bytes per launch of dispatch_kernel at the bench configuration (MessageMesh 4096^2): single-pass metrics, no replay
ARGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dispatch_kernel -s 4 -c 2 --csv --log-file gpurun_out/traffic_4096.csv python bench.py $ARGS > gpurun_out/ncu_traffic.log 2>&1
tail -8 gpurun_out/traffic_4096.csv
# launch list (every kernel with its device time) for the same command
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/launches_4096.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
tail -12 gpurun_out/launches_4096.csv | cut -c1-200
