#!/bin/bash
SSTB200_NVCC_EXTRA="-DSSTB200_TIMING" python -m sst_core_b200.build --force > gpurun_out/build_t.log 2>&1 || tail -20 gpurun_out/build_t.log
python profiles/phase_times.py 2048
python profiles/phase_times.py 4096
python -m sst_core_b200.build --force > /dev/null 2>&1
