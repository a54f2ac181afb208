#!/bin/bash
# 4 GPUs: the C3 strong-scaling bench line at N = 4
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2>gpurun_out/r2_bench_n4.err; echo "bench rc $?"
tail -3 gpurun_out/r2_bench_n4.err | cut -c1-300
