#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are final.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rng.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python profiles/rng_bench.py > gpurun_out/rng_bench.json; python profiles/rng_print.py 2>/dev/null | head -4
