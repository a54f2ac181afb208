#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/build/ubench_tc 2>&1 | head -12 | tee gpurun_out/r2_ubench_issue.txt
export KNN_PROBE_NO_EXACT=1
( SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg16.so timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -4
  SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg16.so KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -4 ) | tee gpurun_out/r2_knn_counters.log
timeout 300 python tools/eig_probe.py 50 100 2>&1 | tail -8 | tee gpurun_out/r2_eig_probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or jacobi" 2>&1 | tail -15 | tee gpurun_out/r2_pytest_eig.log
