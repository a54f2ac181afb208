#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-220
timeout 300 python tools/eig_probe.py 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
timeout 300 python tools/pipeline_probe.py 100000 50 11 > gpurun_out/pipeline_probe.log 2>&1; cat gpurun_out/pipeline_probe.log
