#!/bin/bash
# build variants on the GPU box and time mesh 4096^2 / PHOLD 1024^2 (kernel only)
for mb in 3 5 4; do
  SSTB200_NVCC_EXTRA="-DSSTB200_MIN_BLOCKS=$mb" python -m sst_core_b200.build --force > gpurun_out/build_$mb.log 2>&1
  grep -A2 "dispatch_kernelILb0" gpurun_out/build_$mb.log | grep -E "spill|registers"
  for se in 1280 1536; do
    python - <<PY
import json,subprocess,sys
out=subprocess.run([sys.executable,"bench.py","--steps","20","--warmup","5","--no-cpu-baseline","--no-e2e","--smem-events","$se"],capture_output=True,text=True)
try:
    d=json.loads(out.stdout.strip().splitlines()[-1]); print("minblocks $mb smem_events $se mesh4096",round(d["value"]/1e9,2),"G ev/s",round(d["ms_per_step"],3),"ms")
except Exception as e:
    print("FAILED",out.stdout[-300:],out.stderr[-600:])
PY
  done
done
