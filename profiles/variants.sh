#!/bin/bash
export SSTB200_LIB=$PWD/profiles/micro/libsstb200_racy.so
for i in 1 2; do timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "clock_population" 2>&1 | tail -1; done
