#!/bin/bash
# DRAM bytes and duration per launch of the window kernels at the bench configuration (MessageMesh 4096^2):
# single-pass metrics, no replay.  The launches of one window are window_parallel_kernel, dispatch_kernel (fallback
# list), maintenance_kernel, advance_kernel.
ARGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'window_parallel_kernel|maintenance_kernel|dispatch_kernel|advance_kernel' -s 14 -c 16 --csv --log-file gpurun_out/traffic_4096.csv python bench.py $ARGS > gpurun_out/ncu_traffic.log 2>&1
python profiles/traffic_table.py gpurun_out/traffic_4096.csv
