#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
timeout 300 python tools/knn_probe.py 100000 50 11 1000000 100 21 > gpurun_out/knn_probe.log 2>&1; cat gpurun_out/knn_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv 2>&1 | head -40 | cut -c1-120
