#!/bin/bash
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/knn_launches.csv python tools/knn_probe.py 100000 50 11 > gpurun_out/knn_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/knn_launches.csv | head -20 | cut -c1-140
