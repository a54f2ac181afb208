#!/bin/bash
# round 2, run w: kNN operand preparation with eight lanes per row
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain or mapping" 2>&1 | tail -3 | cut -c1-300
KNN_PROBE_NQ=125000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_launches.csv python tools/knn_probe.py 1000000 100 21 > /dev/null 2>&1
python tools/ncu_times.py gpurun_out/knn_launches.csv 2>&1 | head -12 | cut -c1-120
KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-200
) 2>&1 | tee gpurun_out/r2_w.log
