#!/bin/bash
# round 2, run l: is the C3-shard kNN kernel bound by DRAM (CTAs out of step on a reference set larger than L2)?
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1 KNN_PROBE_NQ=125000
( for pair in 0 1; do for whole in 0 1; do
  echo "== pair=$pair whole=$whole: full, then MMA+TMA only"
  unset SCARF_B200_LIB
  SCF_KNN_PAIR=$pair SCF_KNN_WHOLE=$whole timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
  export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg32.so
  SCF_KNN_PAIR=$pair SCF_KNN_WHOLE=$whole timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
done; done
unset SCARF_B200_LIB
echo "== ncu dram bytes, pair=0 whole=0 / whole=1"
for whole in 0 1; do
SCF_KNN_PAIR=0 SCF_KNN_WHOLE=$whole timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:knn_tc_kernel --csv --log-file gpurun_out/knn_c3_dram_$whole.csv python tools/knn_probe.py 1000000 100 21 > /dev/null 2>&1
grep -i "knn_tc" gpurun_out/knn_c3_dram_$whole.csv | cut -d, -f5,13- | head -12
done
) 2>&1 | tee gpurun_out/r2_l.log
