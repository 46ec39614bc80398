#!/bin/bash
# 2-GPU check of the peer arena + hinted create, and the 1-GPU line with the 100 ns end-to-end leg
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -E "bit-exact|digest|MULTI_GPU_OK|Error|error" | tail -12
SSTB200_CREATE_TIMING=1 timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
grep -E "create\]|peer\]" gpurun_out/bench_${N}gpu.err | tail -40
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu.log").read().strip().splitlines()[-1])
print("mesh $N gpus: value %.4g ms %.4f frac %.3f e2e %.4g (%s) parity %s exchange %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["breakdown"], d["parity"], d["config"]["exchange"]))
PY
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1gpu.log").read().strip().splitlines()[-1])
print("mesh 1 gpu: value %.4g ms %.4f frac %.3f e2e %.4g (%s) parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"], d["parity"]))
PY
tail -3 gpurun_out/bench_1gpu.err
