#!/bin/bash
# round 2, run x: re-rank with candidates pruned by score
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain or mapping or golden" 2>&1 | tail -3 | cut -c1-300
KNN_PROBE_NQ=125000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_launches.csv python tools/knn_probe.py 1000000 100 21 > /dev/null 2>&1
python tools/ncu_times.py gpurun_out/knn_launches.csv 2>&1 | head -6 | cut -c1-120
timeout 600 python tools/knn_fail_probe.py C3 2>&1 | tail -2
timeout 300 python tools/knn_fail_probe.py C2 2>&1 | tail -2
unset KNN_PROBE_NO_EXACT
timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-260
) 2>&1 | tee gpurun_out/r2_x.log
