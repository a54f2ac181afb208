#!/bin/bash
# round 2, run u: launch list of one C3 step on one GPU (what runs next to the kNN tensor kernel)
mkdir -p gpurun_out
( timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload C3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --legs none --no-parity --profiler-range > gpurun_out/bench_ncu_c3.log 2>&1
python tools/ncu_times.py gpurun_out/launches_c3.csv > gpurun_out/launches_c3.txt 2>&1; head -34 gpurun_out/launches_c3.txt | cut -c1-140
) 2>&1 | tee gpurun_out/r2_u.log
