#!/bin/bash
# build variants on the GPU box and time mesh 4096^2 (kernel only)
run() { python - <<PY
import json,subprocess,sys
out=subprocess.run([sys.executable,"bench.py","--steps","20","--warmup","5","--no-cpu-baseline","--no-e2e","--smem-events","$2","--bin-capacity","$3"],capture_output=True,text=True)
try:
    d=json.loads(out.stdout.strip().splitlines()[-1]); print("$1 smem_events $2 cap $3 mesh4096",round(d["value"]/1e9,2),"G ev/s",round(d["ms_per_step"],3),"ms")
except Exception as e:
    print("FAILED",out.stdout[-300:],out.stderr[-600:])
PY
}
SSTB200_NVCC_EXTRA="-DSSTB200_BLOCK=128 -DSSTB200_MIN_BLOCKS=8" python -m sst_core_b200.build --force > gpurun_out/build_a.log 2>&1; grep -A2 "dispatch_kernelILb0" gpurun_out/build_a.log | grep -E "registers"
run "block128/min8" 768 1024
run "block128/min8" 1024 1024
SSTB200_NVCC_EXTRA="-DSSTB200_BLOCK=128 -DSSTB200_MIN_BLOCKS=6" python -m sst_core_b200.build --force > gpurun_out/build_b.log 2>&1; grep -A2 "dispatch_kernelILb0" gpurun_out/build_b.log | grep -E "registers"
run "block128/min6" 768 1024
SSTB200_NVCC_EXTRA="-DSSTB200_BLOCK=512 -DSSTB200_MIN_BLOCKS=2" python -m sst_core_b200.build --force > gpurun_out/build_c.log 2>&1; grep -A2 "dispatch_kernelILb0" gpurun_out/build_c.log | grep -E "registers"
run "block512/min2" 3072 4096
python -m sst_core_b200.build --force > /dev/null 2>&1
