#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --breakdown > gpurun_out/bench_bd.log 2> gpurun_out/bench_bd.err
grep breakdown gpurun_out/bench_bd.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_bd.log') if l.startswith('{')][-1]);print(d['ms_per_step'], d['roofline']['ms_per_launch'])"
