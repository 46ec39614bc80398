#!/bin/bash
# quick GPU check used during kernel work: parity tests, then mesh 4096^2 and PHOLD 1024^2 bench lines (kernel only)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_4096.log 2>gpurun_out/bench_4096.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_4096.log').read().strip().splitlines()[-1]);print('mesh4096',d['value'],d['ms_per_step'],d['roofline']['frac'])"; tail -3 gpurun_out/bench_4096.err
timeout 600 python bench.py --workload phold --x 1024 --y 1024 --steps 20 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_phold.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_phold.log').read().strip().splitlines()[-1]);print('phold1024',d['value'],d['ms_per_step'],d['roofline']['frac'])"
