#!/bin/bash
ARGS="--workload phold --x 1024 --y 1024 --steps 6 --warmup 10 --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dispatch_kernel -s 12 -c 1 -f -o gpurun_out/prof_phold python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
