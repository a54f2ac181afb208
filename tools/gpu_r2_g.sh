#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or jacobi or tridiag or chain or kmeans or symmetris" 2>&1 | tail -15 | cut -c1-250
timeout 300 python tools/eig_probe.py 50 100 2>&1 | tail -3 | tee gpurun_out/r2_eig_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_eig_launches.csv python tools/eig_probe.py 50 100 > /dev/null 2>&1
python tools/ncu_times.py gpurun_out/r2_eig_launches.csv 2>&1 | grep -E "kernel|eig_|jacobi|tridiag" | cut -c1-130 | tee gpurun_out/r2_eig_launches.txt
