#!/bin/bash
# round 2, run q: list updates moved into the waiting time of the epilogue warps
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain or gram or project" 2>&1 | tail -3 | cut -c1-300
KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
KNN_PROBE_NQ=500000 timeout 300 python tools/knn_probe.py 4000000 100 11 2>&1 | tail -1 | cut -c1-200
) 2>&1 | tee gpurun_out/r2_q.log
