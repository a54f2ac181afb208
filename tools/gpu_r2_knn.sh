#!/bin/bash
# kNN kernel check: parity tests that touch the kNN + probe at the BASELINE shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain or mapping" > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_knn.log; tail -5 gpurun_out/r2_pytest_knn.log
timeout 600 python tools/knn_probe.py 5000 25 11 20000 50 11 100000 50 11 100000 100 21 2>&1 | tee gpurun_out/r2_knn_probe.log
KNN_PROBE_NQ=125000 timeout 600 python tools/knn_probe.py 1000000 100 21 2>&1 | tee -a gpurun_out/r2_knn_probe.log
KNN_PROBE_NO_EXACT=1 timeout 600 python tools/knn_probe.py 1000000 100 21 2>&1 | tee -a gpurun_out/r2_knn_probe.log
