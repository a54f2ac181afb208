#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python tools/knn_probe.py 5000 25 11 100000 50 11 200000 100 21 1000000 100 21 > gpurun_out/knn_probe.log 2>&1; cat gpurun_out/knn_probe.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
