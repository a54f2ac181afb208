#!/bin/bash
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -o gpurun_out/knn_tc_c2_full -f python tools/knn_probe.py 100000 50 11 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
