#!/bin/bash
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 1 -o gpurun_out/r2_knn_c2 python tools/knn_probe.py 100000 50 11 > gpurun_out/r2_ncu_knn.log 2>&1
tail -3 gpurun_out/r2_ncu_knn.log
ls -la gpurun_out/*.ncu-rep
