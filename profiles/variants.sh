#!/bin/bash
for v in "-DSSTB200_MESH_PREFETCH=0" "-DSSTB200_MESH_PREFETCH=1" "-DSSTB200_MESH_PREFETCH=2"; do
  SSTB200_NVCC_EXTRA="$v" python -m sst_core_b200.build --force > gpurun_out/build_v.log 2>&1
  ARGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']/1e9,2), 'G ev/s', round(d['ms_per_step'],3), 'ms')"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dispatch_kernel -s 4 -c 1 --csv --log-file gpurun_out/t.csv python bench.py $ARGS > /dev/null 2>&1
  grep dram__ gpurun_out/t.csv | awk -F'","' '{print $13, $15}'
done
python -m sst_core_b200.build --force > /dev/null 2>&1
