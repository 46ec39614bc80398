#!/bin/bash
# one `ncu --set full` capture of dispatch_kernel on MessageMesh 2048^2 (the plain run first, as the recipe demands)
ARGS="--x 2048 --y 2048 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline ${EXTRA}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-window_parallel_kernel} -s ${SKIP:-4} -c 1 -f -o gpurun_out/${1:-prof} python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
