#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-220
timeout 600 python bench.py --steps 3 --warmup 3 --knn-method 1 --gram-mode 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_m1g3.json 2> gpurun_out/bench_m1g3.err; cat gpurun_out/bench_m1g3.json; tail -3 gpurun_out/bench_m1g3.err
