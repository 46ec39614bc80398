#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are final.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_larger_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mesh.log 2> gpurun_out/bench_mesh.err
tail -c 300 gpurun_out/bench_mesh.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_mesh.log').read().strip().splitlines()[-1])
print('mesh', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['match'], d['e2e']['value']/1e9, d['e2e']['breakdown'], d['steady_state']['ms_per_step'], d['steady_state']['value']/1e9)
PY
