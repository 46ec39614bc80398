#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_multi.log
for wl in clockpop phold; do
extra=""; [ $wl = phold ] && extra="--x 1024 --y 1024"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl $extra --steps 20 --warmup 10 --no-e2e > gpurun_out/bench_${wl}_2gpu.log 2> gpurun_out/bench_${wl}_2gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${wl}_2gpu.log') if l.startswith('{')][-1])
    print('$wl 2 GPU:', round(d['value'] / 1e9, 2), 'G/s', round(d['ms_per_step'], 4), 'ms/window  dispatch', round(d['roofline']['ms_per_launch'], 4))
except Exception as e:
    print('$wl FAILED', e); print(open('gpurun_out/bench_${wl}_2gpu.err').read()[-1500:])
PY
done
