#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are final.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
for y in 512 1024 4096; do
timeout 300 python bench.py --x 4096 --y $y --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-steady > gpurun_out/bench_mesh.log 2> gpurun_out/bench_mesh.err
tail -c 300 gpurun_out/bench_mesh.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_mesh.log').read().strip().splitlines()[-1])
print('mesh', d['config']['components'], round(d['value']/1e9,2), 'ms/window', round(d['ms_per_step'],4), 'kernels', round(d['roofline']['ms_per_launch'],4), round(d['roofline']['frac'],4))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'window_parallel_kernel|maintenance_kernel' -s 20 -c 8 --csv --log-file gpurun_out/launch_512.csv python bench.py --x 4096 --y 512 --steps 6 --warmup 10 --no-e2e --no-cpu-baseline --no-steady > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launch_512.csv')) if len(r)>10]
for r in rows[1:]: print(r[4][:40], r[7], r[8], r[-1])
PY
