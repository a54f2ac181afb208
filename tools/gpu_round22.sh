#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/knn_probe.py 100000 50 11 1000000 100 21 2>&1 | tee gpurun_out/knn_probe.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2>gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print(d['ms_per_step'], d['stage_ms']); print(d['roofline']); print(d['e2e'])"; tail -3 gpurun_out/bench_b.err
