#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are final.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
ARGS="--workload phold --x 1024 --y 1024 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
for lib in "" profiles/micro/libsstb200_t256.so profiles/micro/libsstb200_t320.so profiles/micro/libsstb200_t448.so; do
  if [ -n "$lib" ]; then export SSTB200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py $ARGS > gpurun_out/bench_phold.log 2> gpurun_out/bench_phold.err
  tail -c 300 gpurun_out/bench_phold.err
  python - <<'PY'
import json,os
d=json.loads(open('gpurun_out/bench_phold.log').read().strip().splitlines()[-1])
print('phold', os.environ.get('SSTB200_LIB'), round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['digest'])
PY
done
unset SSTB200_LIB
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
