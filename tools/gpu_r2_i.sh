#!/bin/bash
# round 2, run i: new Cholesky / tridiagonal kernels, LOWESS early exit
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or tridiag or jacobi or lowess or hvg or chain or pca" 2>&1 | tail -5 | cut -c1-300
export EIG_PROBE_COV=tools/build/c3_cov.npy EIG_PROBE_N=1000000
SCF_EIG_DEBUG=1 timeout 300 python tools/eig_probe.py 100 2>&1 | grep tridiag | head -2
timeout 300 python tools/eig_probe.py 100 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eig_c3_launches.csv python tools/eig_probe.py 100 > gpurun_out/eig_c3_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/eig_c3_launches.csv > gpurun_out/eig_c3_launches.txt 2>&1; grep -v "sytrd\|laed\|cutlass\|syherk\|larft\|transpose\|stedc" gpurun_out/eig_c3_launches.txt | head -14 | cut -c1-120
unset EIG_PROBE_COV EIG_PROBE_N
SCF_EIG_DEBUG=1 timeout 300 python tools/eig_probe.py 50 2>&1 | grep tridiag | head -1
timeout 300 python tools/eig_probe.py 50 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --legs none --no-parity --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -16 gpurun_out/launches.txt | cut -c1-140
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --legs none --no-parity 2>/dev/null | python -c "
import json,sys
s=sys.stdin.read(); d=json.loads(s[s.index('{\"metric'):].splitlines()[0]); print(d['ms_per_step'], d.get('stage_ms'), d.get('eig'))"
) 2>&1 | tee gpurun_out/r2_i.log
