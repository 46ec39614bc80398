#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "multi_gpu_check exit $?"; tail -3 gpurun_out/multi_gpu_check.log
for g in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 20 --warmup 5 --breakdown > gpurun_out/bench_${g}gpu.log 2> gpurun_out/bench_${g}gpu.err
grep "breakdown\] rank 0" gpurun_out/bench_${g}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${g}gpu.log') if l.startswith('{')][-1])
    print('$g GPU:', round(d['value'] / 1e9, 2), 'G ev/s', round(d['ms_per_step'], 4), 'ms/window  dispatch', round(d['roofline']['ms_per_launch'], 4), 'ms  e2e', round(d['e2e']['value'] / 1e9, 3), 'G ev/s in', round(d['e2e']['seconds'], 3), 's  launches', d['gpu_launches'])
except Exception as e:
    print('$g GPU: FAILED', e); print(open('gpurun_out/bench_${g}gpu.err').read()[-1500:])
PY
done
