#!/bin/bash
# multi-GPU validation + strong-scaling bench lines (run with gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_multi.log
for g in $(seq 1 8); do
  [ $g -le $N ] || continue
  [ $g -eq 1 ] || [ $g -eq 2 ] || [ $g -eq 4 ] || [ $g -eq 8 ] || continue
  if [ $g -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${g}gpu.log 2> gpurun_out/bench_${g}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/bench_${g}gpu.log 2> gpurun_out/bench_${g}gpu.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${g}gpu.log') if l.startswith('{')][-1])
    print('$g GPU:', round(d['value'] / 1e9, 2), 'G ev/s', round(d['ms_per_step'], 4), 'ms/window  dispatch', round(d['roofline']['ms_per_launch'], 4), 'ms  e2e', round(d['e2e']['value'] / 1e9, 3), 'G ev/s in', round(d['e2e']['seconds'], 3), 's  launches', d['gpu_launches'])
except Exception as e:
    print('$g GPU: FAILED', e); print(open('gpurun_out/bench_${g}gpu.err').read()[-1500:])
PY
done
