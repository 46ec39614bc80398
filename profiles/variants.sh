#!/bin/bash
run() { python - <<PY
import json,subprocess,sys
for wl,extra in (("mesh",["--smem-events","$2"]),("phold",["--x","1024","--y","1024","--warmup","10"])):
    out=subprocess.run([sys.executable,"bench.py","--workload",wl,"--steps","20","--warmup","5","--no-cpu-baseline","--no-e2e"]+extra,capture_output=True,text=True)
    try:
        d=json.loads(out.stdout.strip().splitlines()[-1]); print("$1 emax $2",wl,round(d["value"]/1e9,2),"G ev/s",round(d["ms_per_step"],3),"ms")
    except Exception as e:
        print("FAILED",out.stdout[-300:],out.stderr[-600:])
PY
}
for v in "-DSSTB200_MIN_BLOCKS=2" "-DSSTB200_MIN_BLOCKS=3"; do
  SSTB200_NVCC_EXTRA="$v" python -m sst_core_b200.build --force > gpurun_out/build_v.log 2>&1; grep -A2 "dispatch_kernelILb" gpurun_out/build_v.log | grep -E "spill|registers"
  run "$v" 1536
done
python -m sst_core_b200.build --force > /dev/null 2>&1
