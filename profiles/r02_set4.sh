#!/bin/bash
# fused tail kernel: parity tests, then the mesh at full and at per-rank-of-8 size, PHOLD, clock population
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for y in 4096 512; do
timeout 600 python bench.py --x 4096 --y $y --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_mesh_$y.log 2> gpurun_out/bench_mesh_$y.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_mesh_$y.log').read().strip().splitlines()[-1]);print('mesh $y', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['gpu_launches'], d['e2e']['value'], d['e2e']['breakdown'], d['steady_state'] and d['steady_state']['ms_per_step'], d['parity']['match'])"
done
