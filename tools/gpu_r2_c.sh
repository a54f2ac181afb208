#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain or mapping or eig" > gpurun_out/r2_pytest_knn.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_knn.log; tail -12 gpurun_out/r2_pytest_knn.log
timeout 600 python tools/knn_probe.py 100000 50 11 100000 100 21 2>&1 | tee gpurun_out/r2_knn_probe.log
KNN_PROBE_NQ=125000 timeout 600 python tools/knn_probe.py 1000000 100 21 2>&1 | tee -a gpurun_out/r2_knn_probe.log
export KNN_PROBE_NO_EXACT=1
for d in 16 32 160; do
  export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg$d.so
  echo "== SCF_KNN_DEBUG=$d"
  timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -2
  KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -2
done | tee gpurun_out/r2_knn_flags.log
unset SCARF_B200_LIB
timeout 300 python tools/eig_probe.py 50 100 2>&1 | tail -4 | tee gpurun_out/r2_eig_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_eig_launches.csv python tools/eig_probe.py 100 > /dev/null 2>&1
python tools/ncu_times.py gpurun_out/r2_eig_launches.csv 2>&1 | head -24 | cut -c1-130 | tee gpurun_out/r2_eig_launches.txt
