#!/bin/bash
# full GPU check of a tree: parity suite, kNN probe, bench line (with e2e and the CPU baseline), launch list of one step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/knn_probe.py 100000 50 11 1000000 100 21 2>&1 | tee gpurun_out/knn_probe.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2>gpurun_out/bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['ms_per_step'], d['stage_ms']); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks'])"; tail -3 gpurun_out/bench_full.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -24 gpurun_out/launches.txt | cut -c1-140
