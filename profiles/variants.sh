#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are final.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
ARGS="--workload clockpop --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py $ARGS > gpurun_out/bench_clockpop.log 2> gpurun_out/bench_clockpop.err
tail -c 300 gpurun_out/bench_clockpop.err
python - <<'PY'
import json,os
d=json.loads(open('gpurun_out/bench_clockpop.log').read().strip().splitlines()[-1])
print('clockpop', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['digest'])
PY
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:window_clock -s 12 -c 4 --csv --log-file gpurun_out/traffic_clockpop.csv python bench.py --workload clockpop --steps 6 --warmup 10 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/traffic_clockpop.csv')) if len(r)>10]
for r in rows[1:]: print(r[4][:40], r[-3], r[-1])
PY

