#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain or mapping" > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_knn.log; tail -5 gpurun_out/r2_pytest_knn.log
timeout 600 python tools/knn_probe.py 5000 25 11 100000 50 11 100000 100 21 2>&1 | tee gpurun_out/r2_knn_probe.log
KNN_PROBE_NQ=125000 timeout 600 python tools/knn_probe.py 1000000 100 21 2>&1 | tee -a gpurun_out/r2_knn_probe.log
KNN_PROBE_NQ=500000 KNN_PROBE_NO_EXACT=1 timeout 600 python tools/knn_probe.py 4000000 100 11 2>&1 | tee -a gpurun_out/r2_knn_probe.log
export KNN_PROBE_NO_EXACT=1
for d in 8 32 64 96 160; do
  export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg$d.so
  echo "== SCF_KNN_DEBUG=$d"
  timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1
  KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1
done | tee gpurun_out/r2_knn_flags.log
