#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python -c "
import json,sys
for l in open('gpurun_out/bench_${N}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['stage_ms'], d['e2e'])"
tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-300
