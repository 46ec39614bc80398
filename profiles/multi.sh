#!/bin/bash
# multi-GPU measurement set: bash profiles/multi.sh N  (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -E "bit-exact|digest|MULTI_GPU_OK|Error|error" | tail -12
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu.log").read().strip().splitlines()[-1])
print("mesh $N gpus: value %.4g ms %.4f frac %.3f e2e %.4g (%s) parity %s exchange %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["breakdown"], d["parity"]["match"], d["config"]["exchange"]))
PY
tail -3 gpurun_out/bench_${N}gpu.err
for w in clockpop phold; do
  X=""; [ $w = phold ] && X="--x 1024 --y 1024"
  timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload $w $X --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_${w}_${N}gpu.log 2> gpurun_out/bench_${w}_${N}gpu.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_${w}_${N}gpu.log').read().strip().splitlines()[-1]);print('$w $N gpus', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
