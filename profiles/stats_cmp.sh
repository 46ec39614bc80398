#!/bin/bash
for st in 0 1; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --stats $st > gpurun_out/bench_stats$st.log 2>&1
  python -c "
import json; d=json.loads(open('gpurun_out/bench_stats$st.log').read().strip().splitlines()[-1]); print('stats $st', round(d['value']/1e9,2), 'G ev/s', round(d['ms_per_step'],3), 'ms')"
done
