#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json
for l in open('gpurun_out/bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['stage_ms'], d['e2e'])"
tail -2 gpurun_out/bench_2gpu.err
