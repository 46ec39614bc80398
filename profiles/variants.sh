#!/bin/bash
mkdir -p gpurun_out
ARGS="--workload clockpop --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
for v in 1 2 3; do
  timeout 300 python bench.py $ARGS > gpurun_out/bench_clockpop.log 2> gpurun_out/bench_clockpop.err
  python - <<'PY'
import json,os
d=json.loads(open('gpurun_out/bench_clockpop.log').read().strip().splitlines()[-1])
print('clockpop', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['digest'])
PY
done
timeout 600 python -m pytest tests/test_coretest_component.py tests/test_gpu_engine.py tests/test_random_graphs.py tests/test_periodic_stats.py tests/test_gpu_checkpoint.py tests/test_gpu_multirank.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
