#!/bin/bash
# One gpurun call: GPU parity tests, kNN probe, bench (both kNN methods), ncu launch list of one timed step.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/knn_probe.py > gpurun_out/knn_probe.log 2>&1; cat gpurun_out/knn_probe.log
timeout 600 python bench.py --steps 3 --warmup 3 --knn-method 1 > gpurun_out/bench_m1.json 2> gpurun_out/bench_m1.err; cat gpurun_out/bench_m1.json; tail -3 gpurun_out/bench_m1.err
timeout 600 python bench.py --steps 2 --warmup 1 --knn-method 0 --no-e2e --no-cpu-baseline > gpurun_out/bench_m0.json 2> gpurun_out/bench_m0.err; cat gpurun_out/bench_m0.json; tail -3 gpurun_out/bench_m0.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --knn-method 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; cat gpurun_out/launches_summary.txt
