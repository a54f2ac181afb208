#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-220
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
timeout 300 python tools/pipeline_probe.py 100000 50 11 > gpurun_out/pipeline_probe.log 2>&1; tail -3 gpurun_out/pipeline_probe.log
timeout 300 python tools/knn_probe.py 100000 50 11 1000000 100 21 1000000 50 11 > gpurun_out/knn_probe.log 2>&1; cat gpurun_out/knn_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_knn1m.csv -k regex:knn_ python tools/knn_probe.py 1000000 100 21 > /dev/null 2>&1
python tools/ncu_times.py gpurun_out/launches_knn1m.csv 2>&1 | head -12 | cut -c1-120
