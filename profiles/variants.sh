#!/bin/bash
# scratch script for one-off GPU experiments during kernel work (gpurun -- 'bash profiles/variants.sh'); the measurement
# sets that produce the files under profiles/ are r02_set5.sh (one GPU), multi.sh (N GPUs), ncu_full.sh, ncu_traffic.sh,
# timing.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_element_registration.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -5
