#!/bin/bash
# round 2, first call: tcgen05 microbenchmark + the state of the tree on today's box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r2_gpu.txt
timeout 120 tools/build/ubench_tc 2>&1 | tee gpurun_out/r2_ubench_tc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_base.json 2>gpurun_out/r2_bench_c2_base.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_base.json')); print(d['value'], d['ms_per_step'], d['stage_ms']); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['ms_per_launch']); print(d['clocks'])"
tail -3 gpurun_out/r2_bench_c2_base.err
