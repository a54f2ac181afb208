#!/bin/bash
# round 2, run m: rounds + remainder schedule of the kNN kernels; eig small kernels v3
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or tridiag or eig" 2>&1 | tail -4 | cut -c1-300
for pair in 0 1; do
  echo "== pair=$pair: C3/8 full, C3/8 MMA+TMA only, C3 full, C2 full, C4/8 shard"
  unset SCARF_B200_LIB
  SCF_KNN_PAIR=$pair KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
  export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg32.so
  SCF_KNN_PAIR=$pair KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
  unset SCARF_B200_LIB
  SCF_KNN_PAIR=$pair timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
  SCF_KNN_PAIR=$pair timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-200
  SCF_KNN_PAIR=$pair KNN_PROBE_NQ=500000 timeout 300 python tools/knn_probe.py 4000000 100 11 2>&1 | tail -1 | cut -c1-200
done
export EIG_PROBE_COV=tools/build/c3_cov.npy EIG_PROBE_N=1000000
SCF_EIG_DEBUG=1 timeout 300 python tools/eig_probe.py 100 2>&1 | grep tridiag | head -1
timeout 300 python tools/eig_probe.py 100 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eig_c3_launches.csv python tools/eig_probe.py 100 > gpurun_out/eig_c3_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/eig_c3_launches.csv > gpurun_out/eig_c3_launches.txt 2>&1; grep -v "sytrd\|laed\|cutlass\|syherk\|larft\|transpose\|stedc" gpurun_out/eig_c3_launches.txt | head -8 | cut -c1-120
) 2>&1 | tee gpurun_out/r2_m.log
