#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
bash profiles/ncu_traffic.sh > gpurun_out/traffic.out 2>&1; grep "dram__" gpurun_out/traffic.out | cut -d, -f13-15 | head -4
for p in 0 1 2; do
  SSTB200_NVCC_EXTRA="-DSSTB200_MESH_PREFETCH=$p" python -c "
import sst_core_b200.build as b; b.build(force=True)" > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('prefetch $p: ',round(d['value']/1e9,2),'G ev/s',round(d['ms_per_step'],4),'ms', round(d['roofline']['frac'],4))"
done
