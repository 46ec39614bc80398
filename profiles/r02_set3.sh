#!/bin/bash
# after the division purge: parity tests, then PHOLD / clock-population / mesh lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
for w in clockpop phold; do
  X=""; [ $w = phold ] && X="--x 1024 --y 1024"
  timeout 600 python bench.py --workload $w $X --steps 20 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}.log 2> gpurun_out/bench_${w}.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_${w}.log').read().strip().splitlines()[-1]);print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_mesh.log 2> gpurun_out/bench_mesh.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_mesh.log').read().strip().splitlines()[-1]);print('mesh', d['value'], d['ms_per_step'], d['roofline']['frac'], d['steady_state'], d['parity']['match'])"
