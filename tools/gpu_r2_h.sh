#!/bin/bash
# round 2, run h: C3 covariance dump, eigensolver breakdown on it, LOWESS early exit, C2 step launch list
mkdir -p gpurun_out
( timeout 600 python tools/dump_cov.py C3 gpurun_out/c3_cov.npy 2>&1 | tail -2
export EIG_PROBE_COV=gpurun_out/c3_cov.npy EIG_PROBE_N=1000000
SCF_EIG_DEBUG=1 timeout 300 python tools/eig_probe.py 100 2>&1 | grep tridiag | head -3
timeout 300 python tools/eig_probe.py 100 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eig_c3_launches.csv python tools/eig_probe.py 100 > gpurun_out/eig_c3_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/eig_c3_launches.csv > gpurun_out/eig_c3_launches.txt 2>&1; head -30 gpurun_out/eig_c3_launches.txt | cut -c1-120
unset EIG_PROBE_COV EIG_PROBE_N
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or tridiag or lowess or hvg or chain" 2>&1 | tail -3 | cut -c1-250
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --legs none --no-parity --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -24 gpurun_out/launches.txt | cut -c1-140
) 2>&1 | tee gpurun_out/r2_h.log
