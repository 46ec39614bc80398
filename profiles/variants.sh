#!/bin/bash
mkdir -p gpurun_out
for v in "-DSSTB200_PREP=8" "-DSSTB200_PREP=4" "-DSSTB200_PREP=8 -DSSTB200_MESH_AHEAD_LOW=10" "-DSSTB200_PREP=8 -DSSTB200_MESH_PREFETCH=1"; do
  SSTB200_NVCC_EXTRA="$v" python -c "
import sst_core_b200.build as b; b.build(force=True)" > /dev/null 2>&1
  grep -A2 "dispatch_kernelILb0ELb1" sst_core_b200/build_ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '
  python bench.py --steps 20 --warmup 30 --no-e2e --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$v: ',round(d['value']/1e9,2),'G ev/s',round(d['ms_per_step'],4),'ms', round(d['roofline']['frac'],4))"; tail -2 gpurun_out/bench.err
done
