#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are r02_set5.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('mesh',round(d['value']/1e9,2),round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
