#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err; tail -c 1100 gpurun_out/bench_default.json
python bench.py --workload phold --x 1024 --y 1024 --steps 20 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('phold',round(d['value']/1e9,2),round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
